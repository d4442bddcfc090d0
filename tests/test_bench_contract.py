"""bench.py prints exactly one JSON line on stdout with the contract's keys (checked on the CPU reference arm,
which needs no GPU; the GPU arm shares the line builder)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    p = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args], cwd=ROOT, env=e, capture_output=True,
                       text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line_with_contract_keys():
    out = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '0', '--replicas', '8', '--atoms', '64', '--md-steps', '5')
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'higher_is_better', 'scaling',
                'vs_baseline', 'dtype', 'data', 'config', 'cpu_baseline', 'e2e', 'impl'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['steps'] == 1 and d['higher_is_better'] is True
    assert d['value'] > 0 and abs(d['value'] - 1000.0 / d['ms_per_step']) < 1e-6 * d['value']
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0 and d['e2e']['value'] == d['value']
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_reference_arm_other_ranks_exit_silently():
    out = run_bench('--impl', 'reference', '--gpus', '2', '--steps', '1', '--warmup', '0',
                    env={'RANK': '1', 'WORLD_SIZE': '2', 'LOCAL_RANK': '1'})
    assert out.strip() == ''


def test_config4_reference_arm_keeps_the_contract():
    """--workload config4 (BASELINE configs[3], AlanineDipeptideVacuum T-REMD): the same line, its own metric name."""
    out = run_bench('--workload', 'config4', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--replicas', '4',
                    '--md-steps', '20')
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and 'AlanineDipeptideVacuum' in d['metric'] and d['config']['atoms'] == 22
    assert d['config']['replicas'] == 4 and d['config']['md_steps'] == 20 and d['value'] > 0
    assert d['cpu_baseline']['kind'] == 'port' and d['e2e']['value'] == d['value']
