"""bench.py's contract on the CPU: the reference arm (`--impl reference`) prints exactly one JSON line with the keys the
driver reads, on the product arm's metric / unit / config, without loading any of the product's native code; ranks other
than 0 print nothing; the product arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import pathlib
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent.parent


def _run(args, env=None, timeout=600):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, cwd=str(ROOT), env=e, timeout=timeout)


def test_reference_arm_prints_one_contract_line():
    r = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "decoded_info_gbit_per_s" and d["unit"] == "Gbit/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    # the product arm's config keys and values for the headline workload (BASELINE.json configs[1])
    assert d["config"] == {"workload": "wimax_3_4b_n576_65536cw", "code": "N=576 K=432 M=144 nnz=2112", "codewords_per_gpu": 65536,
                           "sigma": 1.0, "max_iter": 40, "early_termination": True}
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_stay_silent():
    r = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"], env={"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"})
    assert r.returncode == 0 and r.stdout.strip() == "", r.stdout + r.stderr[-500:]


def test_reference_arm_loads_nothing_of_the_product():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0'];\n"
            "try:\n    runpy.run_path(%r, run_name='__main__')\nexcept SystemExit:\n    pass\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libldpc_b200.so' not in maps and 'libmyldpc_b200.so' not in maps, 'product library mapped in the reference arm'\n"
            "assert not any(m == 'myldpccppapi_b200' or m.startswith('myldpccppapi_b200.') for m in sys.modules), 'product package imported'\n"
            % str(ROOT / "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(ROOT), timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("a GPU is present")
    r = _run(["--steps", "1", "--warmup", "0", "--no-extras"])
    assert r.returncode != 0 and r.stdout.strip() == ""
    assert "no CPU fallback" in r.stderr or "CUDA" in r.stderr
