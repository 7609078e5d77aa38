"""bench.py contract pieces that run without a GPU: the reference arm (the oracle timed on the host cores) prints one
JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--users', '300',
                          '--items', '500', '--d', '16', '--steps', '1', '--warmup', '1', '--cpu-budget', '0.5'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['metric'] == 'predict_rank_pairs_per_s' and line['unit'] == 'pairs/s'
    assert line['higher_is_better'] is True and line['value'] > 0 and line['steps'] == 1 and line['n_gpus'] == 1
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert line['cpu_baseline']['value'] == line['value'] and line['cpu_baseline']['sample']
    assert line['e2e'] == {'value': line['value'], 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config'] and line['gpu_launches'] == 0


def test_reference_arm_under_torchrun_env_only_rank0_prints():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2', LOCAL_RANK='1')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                          '--users', '64', '--items', '64', '--d', '8', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert out.returncode == 0 and not [l for l in out.stdout.splitlines() if l.startswith('{')]
