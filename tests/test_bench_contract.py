"""bench.py's driver contract, checked on CPU through the reference arm (the product arm needs a GPU): one JSON line
on stdout with the agreed keys; layer FLOP table consistent with SURVEY.md's figure."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_layer_flops_sum_to_the_survey_figure():
    sys.path.insert(0, ROOT)
    import bench
    fl = bench.layer_flops()
    assert len(fl) == 16 and sum(f for _, f in fl) == 301_851_475_968
    assert [n for n, _ in fl][:3] == ["inc.0", "down1.0.0", "down1.0.2"] and fl[-1][0] == "conv1.2"


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "DDIM-50 patches/sec" and d["unit"] == "patches/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "patches/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_traffic_stamp_belongs_to_these_kernels():
    """profiles/traffic.json (the ncu --set full DRAM bytes bench.py quotes as roofline.traffic) was captured on exactly the
    kernel sources in the tree; bench.py drops the number the moment csrc/ changes (comments and whitespace aside)."""
    sys.path.insert(0, ROOT)
    import bench
    tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert tj["csrc_fingerprint"] == bench.csrc_fingerprint()
    assert 15.0 < tj["dram_total_GB"] < 20.0 and len(tj["layers"]) == 16


def test_fingerprint_ignores_comments(tmp_path, monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    src = os.path.join(bench.PKG, "csrc")
    dst = tmp_path / "pkg" / "csrc"
    dst.mkdir(parents=True)
    for name in os.listdir(src):
        text = open(os.path.join(src, name)).read()
        (dst / name).write_text(text)
    monkeypatch.setattr(bench, "PKG", str(tmp_path / "pkg"))
    base = bench.csrc_fingerprint()
    f = dst / "conv_px.cuh"
    f.write_text("// a new comment\n" + f.read_text() + "\n/* another\n one */\n")
    assert bench.csrc_fingerprint() == base
    f.write_text(f.read_text().replace("kPxThreads = 384", "kPxThreads = 256"))
    assert bench.csrc_fingerprint() != base
