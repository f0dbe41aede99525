"""CPU: host logic on either side of the hot path — baseline discovery, batching, leaf jobs, CLI surface."""
import os

import numpy as np
import pytest

from visco_b200.msdata import CORR_TYPES, CORR_TYPES_REVERSE, VisData

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sample():
    return VisData.load(os.path.join(ROOT, "tests", "golden", "sample_ms_kat7.npz"))


def test_sample_bundle(sample):
    assert sample.data.shape == (2160, 16, 4) and sample.data.dtype == np.complex64
    assert sample.corr_types == [9, 10, 11, 12] and sample.antenna_names[:2] == ["ANT-0", "ANT-1"]
    assert CORR_TYPES["XX"] == 9 and CORR_TYPES_REVERSE[12] == "YY"


def test_baseline_discovery_excludes_autocorrelations(sample):
    assert sample.baselines() == [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    v = VisData(data=np.zeros((4, 2, 4), np.complex64), antenna1=[0, 1, 1, 2], antenna2=[0, 0, 2, 1],
                antenna_names=["a", "b", "c"])
    assert v.baselines() == [(0, 1), (1, 2)]                 # (min, max) keys, (0,0) dropped  (compress_ms.py:515-520)
    assert v.baselines(antennas=[0, 2, 1]) == [(0, 2), (0, 1), (2, 1)]
    assert list(sample.baseline_rows(0, 1)) == list(range(0, 2160, 6))
    assert sample.corr_index("YY") == 3 and sample.corr_index(10) == 1
    with pytest.raises(ValueError):
        sample.corr_index("RR")


def test_batching_and_leaf_jobs(sample):
    from visco_b200.compress_ms import _leaf_jobs, batch_baselines
    bl = sample.baselines()
    assert [len(b) for b in batch_baselines(bl, 4)] == [4, 2]
    jobs = list(_leaf_jobs(sample, bl[:2], "XX,YY", False))
    assert [j[0] for j in jobs] == [("ANT-0&ANT-1", "XX"), ("ANT-0&ANT-1", "YY"), ("ANT-0&ANT-2", "XX"), ("ANT-0&ANT-2", "YY")]
    assert jobs[0][1].shape == (360, 16) and jobs[0][2].shape == (360,)
    np.testing.assert_array_equal(jobs[1][1], sample.data[sample.baseline_rows(0, 1)][:, :, 3])
    opt = list(_leaf_jobs(sample, bl[:1], "XX,XY,YX,YY", True))
    assert [j[0][1] for j in opt] == ["diagonals", "offdiagonals"] and opt[0][1].shape == (720, 16)
    np.testing.assert_array_equal(opt[1][1][360:], sample.data[sample.baseline_rows(0, 1)][:, :, 2])
    assert len(opt[0][2]) == 720                              # ROWID tiled twice (compress_ms.py:616)


def test_cli_surface_matches_the_reference():
    from click.testing import CliRunner
    from visco_b200.parser_config import COMPRESS_OPTIONS, DECOMPRESS_OPTIONS, cli
    names = [o[0] for o in COMPRESS_OPTIONS]
    for need in ("ms", "zarrstore", "compressor", "level", "correlation", "correlation_optimized", "fieldid", "ddid",
                 "scan", "column", "outcolumn", "batch_size", "decorrelation", "compressionrank", "antennas", "nworkers",
                 "nthreads", "memory_limit", "use_model_data", "flagestimate", "flagvalue"):
        assert need in names
    abbr = {o[0]: o[1] for o in COMPRESS_OPTIONS}
    assert (abbr["ms"], abbr["zarrstore"], abbr["compressionrank"], abbr["decorrelation"], abbr["batch_size"],
            abbr["correlation_optimized"], abbr["column"]) == ("ms", "zs", "cr", "dec", "bs", "co", "col")
    defaults = {o[0]: o[3] for o in COMPRESS_OPTIONS}
    assert defaults["compressor"] == "zstd" and defaults["level"] == 4 and defaults["batch_size"] == 20
    assert defaults["correlation"] == "XX,YY" and defaults["outcolumn"] == "COMPRESSED_DATA"
    # the reference's four options with its defaults, plus the additive --ngpus (default 1 = the reference's behaviour)
    assert {o[0]: o[3] for o in DECOMPRESS_OPTIONS} == {"zarrstore": None, "ms": "decompressed.ms",
                                                        "column": "COMPRESSED_DATA", "batch_size": 50, "ngpus": 1}
    assert defaults["ngpus"] == 1 and abbr["ngpus"] == "ng"
    r = CliRunner()
    out = r.invoke(cli, ["--help"])
    assert out.exit_code == 0 and "compressms" in out.output and "decompressms" in out.output
    out = r.invoke(cli, ["compressms", "--help"])
    assert out.exit_code == 0 and "-cr, --compressionrank" in out.output and "--correlation-optimized" in out.output
    out = r.invoke(cli, ["decompressms", "--help"])
    assert out.exit_code == 0 and "-zs, --zarrstore" in out.output


def test_driver_errors_before_any_gpu_work(tmp_path):
    from visco_b200.compress_ms import compress_full_ms
    kw = dict(zarr_path=str(tmp_path / "z"), consolidated=True, chunk_size_row=100, overwrite=True, compressor="zstd",
              level=3, nworkers=1, nthreads=1, memory_limit="1GB", direct_to_workers=False, correlation="XX,YY",
              correlation_optimized=False, fieldid=0, ddid=0, scan=1, column="DATA", outcolumn="COMPRESSED_DATA",
              batch_size=10)
    with pytest.raises(ValueError):
        compress_full_ms(ms_path=str(tmp_path / "missing.ms"), **kw)                    # reference :876-877
    bundle = os.path.join(ROOT, "tests", "golden", "sample_ms_kat7.npz")
    with pytest.raises(ValueError):
        compress_full_ms(ms_path=bundle, **{**kw, "compressor": "lzma"})                 # reference :51
    with pytest.raises(NotImplementedError):
        compress_full_ms(ms_path=bundle, flag_estimate=True, **kw)                       # griddata estimator: out of scope
    with pytest.raises(ValueError):
        compress_full_ms(ms_path=bundle, flagvalue="not-a-number", **kw)                 # reference :549-558
