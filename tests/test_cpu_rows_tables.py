"""The table builder, ring-entry grouping and byte-ring placement of the row-streaming RoIAlign kernel
(csrc/roi_align_fwd_rows.cu), transcribed lane by lane in tests/rows_tables_emulation.py, against the C
oracle -- the part of that kernel that can be checked without a GPU (the GPU parity tests are
tests/test_gpu_roi_align_rows.py)."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu():
    spec = importlib.util.spec_from_file_location("rows_tables_emulation", os.path.join(ROOT, "tests", "rows_tables_emulation.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_tables_and_consumer_arithmetic_match_the_oracle(emu, monkeypatch, capsys):
    monkeypatch.setattr("sys.argv", ["rows_tables_emulation.py", "40"])
    emu.main()          # asserts rtol 1e-5 per RoI, identical FPN levels, edge RoIs, wide single-level RoIs
    out = capsys.readouterr().out
    assert out.count("ok:") == 2


def test_ring_planner_never_overwrites_a_live_entry(emu):
    emu.check_ring_planner(n_rois=1500, seed=11)
