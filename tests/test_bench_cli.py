"""CPU: bench.py's reference arm prints one JSON line with the contract's keys (the GPU arm
is exercised on the GPU box by the driver and by scripts/full_check.sh)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0", "--genome-mb", "2", "--variants", "100"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "read_bases_kmer_probed_per_sec"
    assert d["unit"] == "bases/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "bases/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None
    assert d["config"]["step_fraction"] == 1.0  # small workloads run the whole step


def test_reference_arm_does_not_load_the_product(tmp_path):
    """The reference arm must not map libdkb.so (nor import the package that binds it): its
    numbers are the oracle's alone."""
    code = (
        "import sys, runpy\n"
        "sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--genome-mb', '1', '--variants', '50']\n"
        "try:\n"
        "    runpy.run_path(%r, run_name='__main__')\n"
        "except SystemExit:\n"
        "    pass\n"
        "maps = open('/proc/self/maps').read()\n"
        "assert 'libdkb' not in maps, 'libdkb.so is mapped'\n"
        "assert 'libdnk_oracle' in maps\n"
        "assert not any(m == 'denovo_kmer_b200' or m.startswith('denovo_kmer_b200.') for m in sys.modules), 'package imported'\n"
        "sys.stderr.write('CLEAN\\n')\n" % os.path.join(ROOT, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "CLEAN" in out.stderr, out.stderr[-2000:]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
