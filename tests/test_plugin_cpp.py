"""The C++ drop-in (ndt_2d_b200/plugin: ndt_2d_b200::ScanMatcherNDT behind the reference's
abstract ndt_2d::ScanMatcher, and ndt_2d_b200::ParticleFilter) driven next to the
reference's own plugin by tests/cpp/plugin_parity.cpp.  The binary embeds the reference's
sources compiled in place, so it is built into oracle/_ref/ where /root/reference exists
and travels to the GPU box from there."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
BIN = ROOT / "oracle" / "_ref" / "plugin_parity"


def _need_binary():
    if not BIN.exists():
        pytest.skip("oracle/_ref/plugin_parity not built (reference tree not available here)")


def test_plugin_sources_follow_reference_interface():
    """Every virtual of ndt_2d::ScanMatcher (scan_matcher.hpp:53-90) is overridden, the six
    parameters keep the reference's names, and plugins.xml keeps its lookup name."""
    hpp = (ROOT / "ndt_2d_b200/plugin/include/ndt_2d_b200/scan_matcher_ndt.hpp").read_text()
    cpp = (ROOT / "ndt_2d_b200/plugin/src/scan_matcher_ndt.cpp").read_text()
    for method in ("initialize", "addScans", "matchScan", "scoreScan", "scorePoints", "reset"):
        assert f" {method}(" in hpp and f"ScanMatcherNDT::{method}(" in cpp
    assert hpp.count("override") >= 6
    for name in ("ndt_resolution", "search_angular_resolution", "search_angular_size",
                 "search_linear_resolution", "search_linear_size", "laser_max_beams"):
        assert f'".{name}"' in cpp
    assert "PLUGINLIB_EXPORT_CLASS(ndt_2d_b200::ScanMatcherNDT, ndt_2d::ScanMatcher)" in cpp
    xml = (ROOT / "ndt_2d_b200/plugin/plugins.xml").read_text()
    assert 'name="ndt_2d::ScanMatcherNDT"' in xml and 'base_class_type="ndt_2d::ScanMatcher"' in xml


def test_plugin_fails_loudly_without_gpu():
    from ndt_2d_b200 import lib
    _need_binary()
    if lib.ndt2d_device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([str(BIN), "--no-gpu"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_plugin_parity_with_reference_plugin(gpu):
    _need_binary()
    r = subprocess.run([str(BIN)], capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert '"plugin_parity": "ok"' in r.stdout
