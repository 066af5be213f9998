"""Repository contracts that need no GPU: the product never touches the oracle, bench.py's CPU arm prints exactly one
JSON line with the agreed keys, and every CUDA entry point of the header is bound by the ctypes table."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _py_files(folder):
    for base, _, files in os.walk(os.path.join(ROOT, folder)):
        for f in files:
            if f.endswith('.py'):
                yield os.path.join(base, f)


def test_product_package_never_imports_the_oracle():
    for path in _py_files('gnnome_b200'):
        text = open(path).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), path
        assert '/root/reference' not in text, path


def test_oracle_is_only_used_by_tests_smoke_and_bench():
    users = []
    for folder in ('gnnome_b200', 'tools'):
        for path in _py_files(folder):
            if re.search(r'^\s*(from|import)\s+oracle\b', open(path).read(), flags=re.M):
                users.append(os.path.relpath(path, ROOT))
    # diagnostics that use the oracle as their checker live under tests/diag/, not in the package or in tools/
    assert users == [], users


def test_bench_reference_arm_prints_one_json_line():
    env = dict(os.environ, GNB_BENCH_CPU_SAMPLE='2000,12000')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                          '--warmup', '0', '--workload', 'small'], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'edges/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config']
