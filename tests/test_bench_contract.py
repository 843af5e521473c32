"""CPU: the output contract of bench.py's reference arm (one JSON line on stdout, the keys the driver reads), and the
forwarding of the drop-in config_io_module."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_json_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "atom-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0
    assert d["metric"].startswith("LJ atom-steps/sec") and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("c5")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_dropin_config_io_forwards_other_names(tmp_path):
    """read_cnf_atoms / write_cnf_atoms come from this repository; every other name resolves to the next
    config_io_module.py on sys.path (the reference's), so the molecular programs keep working."""
    ref = tmp_path / "ref"
    ref.mkdir()
    (ref / "config_io_module.py").write_text("def read_cnf_mols(filename, with_v=False, quaternions=False):\n"
                                             "    return 'reference read_cnf_mols', filename\n")
    code = ("import sys\n"
            "sys.path.insert(0, %r)\n"
            "sys.path.insert(0, %r)\n"
            "import config_io_module as c\n"
            "assert c.read_cnf_atoms.__module__ == 'examples_b200.config_io_module', c.read_cnf_atoms.__module__\n"
            "assert c.read_cnf_mols('x') == ('reference read_cnf_mols', 'x')\n"
            "try:\n"
            "    c.no_such_name\n"
            "except AttributeError:\n"
            "    print('ok')\n") % (str(ref), os.path.join(ROOT, "examples_b200", "dropin"))
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and p.stdout.strip() == "ok", p.stderr[-2000:]
