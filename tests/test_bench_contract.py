"""bench.py contract that can be checked without a GPU: the reference arm prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '0', '--workload', 'small', '--cpu-seconds', '1'],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j['impl'] == 'reference' and j['metric'] == 'knn_pair_dists_per_s' and j['unit'] == 'pair-dists/s'
    assert j['higher_is_better'] is True and j['vs_baseline'] is None and j['data'] == 'synthetic'
    assert j['value'] > 0 and j['steps'] == 1 and j['gpu_launches'] == 0
    from oracle import ref_loader
    assert j['cpu_baseline']['kind'] == ('reference' if ref_loader.available() else 'port')
    assert j['cpu_baseline']['cores'] == os.cpu_count()
    assert j['cpu_baseline']['value'] == j['value'] == j['e2e']['value']
    assert j['e2e']['h2d_bytes_per_step'] == 0 and j['e2e']['d2h_bytes_per_step'] == 0
    # like for like: the config object is the one our arm prints for the same workload, and the value is for the FULL
    # configuration (pairs of the full problem / seconds per logE at full size), extrapolated from the sample
    import bench
    assert j['config'] == bench.config_for('small', 1)
    n = j['config']['n']
    assert abs(j['value'] - n * n / j['s_per_logE']) < 1e-6 * j['value']
    assert j['extrapolated'] is True and j['sample_rows'] < n and 1.0 <= j['used_exponent'] <= 1.3


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2'],
                       capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ''
