"""CPU tests (no GPU) of bench.py's contract: the reference arm runs on host cores alone and
prints exactly one JSON line with the agreed keys; our arm refuses to run without a CUDA device
(there is no CPU fallback to time)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_reference_arm_prints_one_json_line(tmp_path):
    env = dict(os.environ, PPB200_CACHE=str(tmp_path))
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--reads", "30000",
                          "--chunk", "2000", "--steps", "2", "--warmup", "1"], env=env, check=True,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True
    assert d["metric"] == "DecompressAll uncompressed GB/s" and d["value"] > 0 and d["n_gpus"] == 1
    assert d["config"]["records"] == 30000 and "workload" in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["gpu_launches"] == 0 and d["scaling"] == "weak" and d["dtype"] == "u8"


def test_reference_arm_other_ranks_exit_quietly(tmp_path):
    env = dict(os.environ, PPB200_CACHE=str(tmp_path), RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--reads", "1000", "--steps", "1", "--warmup", "0"], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_our_arm_fails_loudly_without_cuda(tmp_path):
    if _have_gpu():
        import pytest
        pytest.skip("checks the no-device behaviour")
    env = dict(os.environ, PPB200_CACHE=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--reads", "1000", "--steps", "1"], env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_reference_arm_never_loads_the_product(tmp_path):
    """The reference arm's corpus index comes from the oracle's own CreateIndex: the process must not
    map libppb200.so nor import the product package (the arm was 'tainted')."""
    code = (
        "import sys, os, runpy\n"
        f"sys.argv = ['bench.py', '--impl', 'reference', '--reads', '20000', '--chunk', '2000', '--steps', '1', '--warmup', '1']\n"
        "try:\n"
        f"    runpy.run_path({os.path.join(ROOT, 'bench.py')!r}, run_name='__main__')\n"
        "except SystemExit:\n"
        "    pass\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libppb200' not in maps, 'product library mapped in the reference arm'\n"
        "assert 'libpporacle' in maps\n"
        "assert not any(m.startswith('parallelparsing_b200') for m in sys.modules), 'product package imported'\n"
    )
    env = dict(os.environ, PPB200_CACHE=str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["warmup"] == 1 and d["steps"] == 1


def test_reference_arm_multi_gpu_file_is_n_times_larger(tmp_path):
    """--gpus N: ONE file of N x reads reads (sharded generation above 10 M reads is exercised with a
    small shard size), the same `config` our arm prints."""
    code = (
        "import sys, os, runpy\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "bench.SHARD_READS = 4000\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--gpus', '4', '--reads', '3000', '--chunk', '500', '--steps', '1', '--warmup', '0']\n"
        "bench.main()\n"
    )
    env = dict(os.environ, PPB200_CACHE=str(tmp_path))
    r = subprocess.run([sys.executable, "-c", code], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][0])
    assert d["n_gpus"] == 4 and d["config"]["file_reads"] == 12000 and d["config"]["records"] >= 12000
    assert "4 contiguous chunk ranges" in d["config"]["partition"] and "shards of 4000 reads" in d["config"]["workload"]
    assert set(d["config"]) == {"workload", "file_reads", "chunks", "records", "uncompressed_bytes", "compressed_bytes",
                                "partition", "l2"}
