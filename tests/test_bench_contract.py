"""CPU checks of what bench.py and the tools rely on: the committed ncu figures behind `roofline`, the workload table, the
reference arm (oracle/_ref on the host cores) printing the contract's JSON line, and that every tool still parses."""
import glob
import json
import os
import py_compile
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_traffic_json_has_what_the_roofline_reads():
    t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    for kernel in ("resize_level_kernel", "fast_tiles_kernel", "blur_all_kernel"):
        assert kernel in t, kernel
    f = t["fast_tiles_kernel"]
    for key in ("dram_bytes_per_launch", "warp_inst_per_launch", "alu_warp_inst_per_launch", "frames_per_pass", "alu_pipe_pct",
                "issue_active_pct"):
        assert key in f and f[key] > 0, key
    # the capture is of the dominant kernel of one 512-frame pass: more DRAM traffic than the algorithmic 950 532 B per frame,
    # far fewer than 1000 warp-instructions per pixel
    assert f["dram_bytes_per_launch"] / f["frames_per_pass"] > 950532
    assert 0 < f["alu_warp_inst_per_launch"] < f["warp_inst_per_launch"] < 1000.0 * 950532 * f["frames_per_pass"] / 32


def test_workloads_and_algorithmic_bytes():
    import bench
    assert set(bench.WORKLOADS) >= {"C3_tum_640x480_1000kp_8lv", "C2_euroc_752x480_1000kp_8lv", "C5_1080p_4000kp_12lv"}
    cfg = bench.config_of("C3_tum_640x480_1000kp_8lv")
    assert cfg["workload"] == "C3_tum_640x480_1000kp_8lv" and cfg["width"] == 640 and cfg["height"] == 480
    frames = bench.make_frames(3, 64, 48)
    assert frames.shape == (3, 48, 64) and frames.dtype.name == "uint8"
    assert (frames[0] != frames[1]).any()


def test_tools_parse():
    for p in glob.glob(os.path.join(ROOT, "tools", "*.py")) + glob.glob(os.path.join(ROOT, "tests", "tools", "*.py")):
        py_compile.compile(p, doraise=True)
    for p in glob.glob(os.path.join(ROOT, "tools", "*.sh")):
        subprocess.check_call(["bash", "-n", p])


def test_reference_arm_prints_the_contract_line():
    from oracle import ref_binding
    if not ref_binding.available():
        pytest.skip("oracle/_ref is not built (no /root/reference in this environment)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=dict(os.environ, LD_DEBUG="libs"))  # the loader logs every library
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0 and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["config"]["workload"] == "C3_tum_640x480_1000kp_8lv"
    # the arm must not map the product library, and it does run the reference build
    assert "libsdorb.so" not in out.stderr
    assert "libsdorb_ref" in out.stderr
