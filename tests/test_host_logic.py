"""CPU tests of the host-side mirror: VideoMode table and lookups (the reference keeps these in
Julia host code, src/VideoConfigurations.jl), zoom_autocorr slicing, lag/line conversions."""
import numpy as np

import tempestsdr_b200 as tsdr


def test_video_mode_table_like_reference_runtests():
    # test/runtests.jl:29-51: a Dict{String,VideoMode} with more than 10 entries, every entry findable
    d = tsdr.allVideoConfigurations
    assert isinstance(d, dict) and len(d) == 80 and len(d) > 10
    assert all(isinstance(k, str) and isinstance(v, tsdr.VideoMode) for k, v in d.items())
    for name, cfg in d.items():
        found = tsdr.find_closest_configuration(cfg.height, cfg.refresh)
        assert any(v == cfg for v in found.values()), name
        assert tsdr.find_configuration(cfg) is not None
    assert sorted(set(v.refresh for v in d.values())) == [25, 30, 43, 56, 60, 65, 70, 72, 75, 76, 85, 100, 120]


def test_find_closest_follows_the_code_not_the_docstring():
    # src/VideoConfigurations.jl:114-115 claims "1024x600 @ 60 Hz"; the code picks the 60 Hz entry whose HEIGHT is nearest
    got = tsdr.find_closest_configuration(1280, 60)
    assert list(got) == ["1600x1200 @ 60Hz"] and tsdr.dict2video(got) == tsdr.VideoMode(2160, 1250, 60)
    assert tsdr.find_configuration(tsdr.VideoMode(2592, 1242, 60)) == "1920x1200 @ 60Hz"
    assert tsdr.find_configuration(tsdr.VideoMode(1, 2, 3)) is None
    assert list(tsdr.find_closest_configuration(1589, 60.14)) == ["2048x1536 @ 60Hz"]  # docs/src/gui.md:29 known answer
    both = tsdr.find_closest_configuration(795, 60)  # two 60 Hz modes share height 795: both are returned
    assert sorted(both) == ["1280x768 @ 60 Hz", "1368x768 @ 60 Hz"]


def test_video_mode_means_total_raster():
    m = tsdr.allVideoConfigurations["1920x1080 @ 60Hz"]
    assert (m.width, m.height, m.refresh) == (2576, 1125, 60.0)
    assert tsdr.getImageDuration(m, 20e6) == 333333


def test_zoom_and_conversions():
    g = np.arange(1, 2_000_001, dtype=np.float32)
    rates, sl = tsdr.zoom_autocorr(g, 20e6, rate_min=50, rate_max=90)
    assert sl[0] == 222222 and sl[-1] == 400000 and rates.size == sl.size
    assert tsdr.delay2yt(1 / (60 * 1125), 60) == 1125 and tsdr.yt2index(1125, 20e6, 60) == 296
    assert abs(tsdr.yt2delay(1125, 60) - 1 / 67500) < 1e-18
    assert tsdr.RENDERING_SIZE == (600, 800)


def test_spectrum_axes_and_contrast_helpers():
    # host arithmetic of the GetSpectrum.jl wrappers (src/GetSpectrum.jl:26,46,63-64) and of the search score
    from tempestsdr_b200 import api
    N, fs = 10, 4.0
    assert np.allclose(api._freq_axis(N, fs), [((k / N) - 0.5) * fs for k in range(N)])
    assert api._freq_axis(1024, 1.0)[512] == 0.0 and api._freq_axis(1024, 1.0)[0] == -0.5
    # flat projection c = m: Sigma = n m; the reference's beta for it is (m/2 + m/2)^2 = m^2 -> contrast 1
    n, m = 800, 3.5
    assert np.allclose(api.blanking_contrast([m * m], [n * m], n), 1.0)
    c = api.blanking_contrast([4.0, 0.0], [10.0, 0.0], 5)       # mean 2 -> 1.0 ; empty frame -> nan, not an exception
    assert c[0] == 1.0 and np.isnan(c[1])


def test_shard_helpers_cover_every_hypothesis_once():
    from tempestsdr_b200 import parallel
    for n, w in [(13, 8), (80, 8), (5, 2), (3, 4)]:
        seen = sorted(sum((parallel.shard_round_robin(n, w, r) for r in range(w)), []))
        assert seen == list(range(n))
        blocks = [parallel.shard_contiguous(n, w, r) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n and all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))


def test_bench_workloads_follow_baseline_configs():
    """bench.py's workload table is BASELINE.json's configs[1..4]: sample rates, total rasters,
    buffer sizes; the headline default is the north-star configuration (cfg 3)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    w = bench.WORKLOADS
    assert (w["cfg2"]["Fs"], w["cfg2"]["x_t"], w["cfg2"]["y_t"], w["cfg2"]["fv"]) == (20e6, 2576, 1125, 60.0)
    assert (w["cfg3"]["Fs"], w["cfg3"]["x_t"], w["cfg3"]["y_t"], w["cfg3"]["fv"]) == (200e6, 2720, 1481, 60.0)
    assert (w["cfg5"]["Fs"], w["cfg5"]["x_t"], w["cfg5"]["y_t"], w["cfg5"]["fv"], w["cfg5"]["total_frames"]) == (200e6, 4400, 2250, 30.0, 1000)
    assert w["cfg4"]["n_ech"] == 1 << 26 and w["cfg3"]["n_ech"] == 10 ** 8
    src = open(os.path.join(root, "bench.py")).read()
    assert 'default="cfg3"' in src and '"cfg5_fullres"' in src          # headline workload; the full-resolution leg is selectable
    assert "TSDR_BENCH_EXTRAS_DEADLINE" in src                          # the collective legs run under a watchdog
