"""bench.py's pieces that need no GPU: the committed ncu capture the roofline block is read from, and the configuration table."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_roofline_source_is_committed_and_consistent():
    b = load_bench()
    rows = b.ncu_sweep_rows()                      # raises SystemExit when profiles/r02_ncu_raw.csv is missing
    assert len(rows) == 2
    algorithmic = 4.0 * (1920 - 128) * 1080 * 128 * b.PROFILE_PAIRS      # bytes per sweep launch, SURVEY.md section 8d
    for r in rows:
        assert 0.98 < r["dram_bytes"] / algorithmic < 1.15, r          # DRAM traffic = algorithmic bytes: no wasted re-reads
        assert 30 < r["alu_pct"] < 100 and 0.005 < r["seconds"] < 0.1


def test_configs_cover_baseline_json():
    b = load_bench()
    assert set(b.CONFIGS) == {"1080p128", "720p64", "bm720p", "ref_hh4", "4k256_pipeline", "stream10k"}
    for name, c in b.CONFIGS.items():
        assert c["kind"] in ("sgbm", "bm", "pipeline", "stream") and c["W"] > c["D"] and c["batch"] >= 1, name
    assert b.CONFIGS["1080p128"]["batch"] == 64 and b.CONFIGS["1080p128"]["mode"] == 1
    assert b.CONFIGS["stream10k"]["frames"] == 10000
