"""CPU-side check (-m "not gpu") of the bench.py contract: the reference arm runs without a GPU and prints one JSON line
with the keys the driver reads; the algorithmic byte / flop counts match SURVEY.md 8(d)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    res = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '3', '--warmup', '1',
                          '--batch', '256'], capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'states/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['gpu_launches'] == 0
    from oracle import make_ref
    assert d['cpu_baseline']['kind'] == ('reference' if make_ref.available() else 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['cpu_baseline']['cores_available'] >= d['cpu_baseline']['cores']
    assert d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0
    assert 'workload' in d['config'] and 'input_pool_mb' in d['config']      # the same config keys in both arms


def test_algorithmic_counts_match_the_survey():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.algorithmic('graph', 5) == (904, 72704)          # 904 B, 36 352 MAC
    assert bench.algorithmic('value', 5) == (140, 2 * 50676)
    assert bench.algorithmic('statepred', 5)[0] == 236
    assert bench.algorithmic('value', 10) == (240, 2 * 86036)
    assert bench.algorithmic('value', 20) == (440, 2 * 171156)
