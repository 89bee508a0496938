"""The drop-in CLI (csrc/main.cpp -> bin/mipb200_main): flag parsing like boost::program_options (unique prefixes,
--opt=value / --opt value, short -f/-s/-o/-l), the parameter echo, CSV errors, and -- on the GPU -- the cost log format."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _cli(mip):
    if not os.path.exists(mip.CLI_PATH):
        mip.build()
    return mip.CLI_PATH


def _run(mip, *args, cwd=None):
    return subprocess.run([_cli(mip), *args], capture_output=True, text=True, cwd=cwd, timeout=600)


def test_help_and_missing_required(mip):
    r = _run(mip, "--help")
    assert r.returncode == 1 and "--FramesToBeEncoded" in r.stdout and "--KernelIdx" in r.stdout
    r = _run(mip, "-s", "256x128")
    assert r.returncode == 1
    assert "[!] ERROR: FramesToBeEncoded not set." in r.stdout and "[!] ERROR: Input original frames not set." in r.stdout
    assert "Exiting after finding errors in input parameters" in r.stdout


def test_prefix_matching_and_echo(mip, tmp_path):
    r = _run(mip, "--Frames=1", "--Res", "250x128", "--Orig", "nofile.csv", "--Device=0")
    assert "-=-= INPUT PARAMETERS =-=-" in r.stdout and "FramesToBeEncoded=1" in r.stdout and "Device Index=0" in r.stdout
    assert "OutputPreffix log file not set" in r.stdout
    assert "[!] ERROR: Unsupported resolution 250x128" in r.stdout and r.returncode == 0      # the reference returns 0 here (main.cpp:301-309)
    r = _run(mip, "--F=1", "-s", "256x128", "-o", "x.csv")      # --F is ambiguous: FramesToBeEncoded / FilterType
    assert r.returncode == 1 and "ambiguous" in r.stderr
    r = _run(mip, "--Bogus=1")
    assert r.returncode == 1 and "unrecognised option" in r.stderr


def test_filter_whitelist_and_runtime_alt_switch(mip):
    r = _run(mip, "-f", "1", "-s", "256x128", "-o", "x.csv", "--UseAlternativeSamples=1", "--Filter=filterFrame_3d", "--KernelIdx=1")
    assert "FilterType=filterFrame_3d" in r.stdout and "KernelIdx=1" in r.stdout
    assert "[!] ERROR: Filter type filterFrame_3d not supported" in r.stdout and r.returncode == 0   # exit(0), main.cpp:76
    r = _run(mip, "-f", "1", "-s", "256x128", "-o", "x.csv", "--UseAlternativeSamples=1")
    # like the reference, the whitelist check (main.cpp:71-77) fires before the error count is looked at (:79-82)
    assert "[!] ERROR: Filter not set." in r.stdout and "Filter type  not supported" in r.stdout and r.returncode == 0


def test_bad_resolution_string_and_missing_file(mip):
    r = _run(mip, "-f", "1", "-s", "1920", "-o", "x.csv")
    assert 'Input resolution "1920" not set properly' in r.stdout
    r = _run(mip, "-f", "1", "-s", "256x128", "-o", "/nonexistent/x.csv")
    assert r.returncode == 1 and "error while opening samples files" in r.stderr


def test_short_csv_is_an_error(mip, tmp_path):
    from mipb200 import frames
    p = tmp_path / "short.csv"
    frames.write_csv(str(p), [frames.noise_frame(256, 100, 1)])       # 100 lines, 128 needed
    r = _run(mip, "-f", "1", "-s", "256x128", "-o", str(p))
    assert r.returncode == 1 and "holds 100 lines, need 128" in r.stderr


@pytest.mark.gpu
def test_cli_log_matches_oracle(mip, oracle, tmp_path):
    """2 frames 256x184 through the real CLI/CSV; frame 0's log rows == oracle in the reference's row order and format."""
    from mipb200 import frames, tables as T
    fs = [frames.noise_frame(256, 184, 40), frames.natural_frame(256, 184, 41)]
    csv = tmp_path / "in.csv"
    frames.write_csv(str(csv), fs)
    np.testing.assert_array_equal(frames.read_csv(str(csv), 256, 184, 2), np.stack(fs))
    for extra, ft, kidx in (([], 0, 0), (["--UseAlternativeSamples=1", "--Filter=filterFrame_2d_float_5x5_quarterCtu", "--KernelIdx=2"], 8, 2)):
        prefix = tmp_path / f"log{ft}"
        r = _run(mip, "-f", "2", "-s", "256x184", "-o", str(csv), "-l", str(prefix), *extra)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "Current frame 0" in r.stdout and "Current frame 1" in r.stdout
        assert "Elapsed time (ms) from writing samples to reading distortion (2x)," in r.stdout
        lines = open(str(prefix) + ".csv").read().splitlines()
        assert lines[0] == "CTU,cuSizeName,W,H,CU,X,Y,Mode,SAD,SATD,minSadHad"
        cost, sad, satd = oracle.run_frame(fs[0], ft, kidx, want_sad_satd=True)
        assert len(lines) - 1 == cost.size
        want = []
        for ctu in range(cost.shape[0]):
            cx, cy = 128 * (ctu % 2), 128 * (ctu // 2)
            for t in T.TYPES:
                for cu in range(t.n):
                    x, y = t.pos(cu)
                    for m in range(t.modes):
                        i = T.COST_OFFSETS[t.idx] + cu * t.modes + m
                        want.append(f"{ctu},{t.name},{t.w},{t.h},{cu},{cx + x},{cy + y},{m},{sad[ctu, i]},{satd[ctu, i]},{cost[ctu, i]}")
        assert lines[1:] == want


@pytest.mark.gpu
def test_cli_all_frames_and_two_gpu_workers(mip, oracle, tmp_path):
    """--AllFrames adds a POC column; --NumGpus shards frames (skipped when the box has a single GPU)."""
    import torch
    from mipb200 import frames
    fs = [frames.noise_frame(128, 128, 60 + i) for i in range(3)]
    csv = tmp_path / "in.csv"
    frames.write_csv(str(csv), fs)
    g = 2 if torch.cuda.device_count() >= 2 else 1
    r = _run(mip, "-f", "3", "-s", "128x128", "-o", str(csv), "-l", str(tmp_path / "all"), "--AllFrames", "--Compat", f"--NumGpus={g}")
    assert r.returncode == 0, r.stdout + r.stderr
    lines = open(str(tmp_path / "all") + ".csv").read().splitlines()
    assert lines[0].startswith("POC,CTU,") and len(lines) - 1 == 3 * 97840
    for poc in range(3):
        cost = oracle.run_frame(fs[poc])
        rows = lines[1 + poc * 97840: 1 + (poc + 1) * 97840]
        assert all(row.startswith(f"{poc},0,") for row in rows[:5])
        got = np.array([int(row.rsplit(",", 1)[1]) for row in rows], dtype=np.int32)
        assert np.array_equal(got, cost[0])
        assert rows[0].split(",")[-3:-1] == ["0", "0"]       # --Compat: SAD/SATD columns print 0 like the reference build


def test_unknown_input_format(mip, tmp_path):
    p = tmp_path / "x.bin"
    p.write_bytes(b"\0" * 16)
    r = _run(mip, "-f", "1", "-s", "128x4", "-o", str(p), "--InputFormat=png")
    assert r.returncode == 1 and "InputFormat png not supported" in r.stdout


@pytest.mark.gpu
def test_cli_binary_input_and_decisions_log(mip, oracle, tmp_path):
    """Raw u16 and 8-bit yuv420p inputs, --NoLog --DecisionsLog: per-CU best mode / cost of every frame == oracle."""
    from mipb200 import frames, tables as T
    fs = [frames.natural_frame(256, 128, 80 + i) for i in range(3)]
    raw = tmp_path / "in.u16"
    np.stack(fs).astype("<u2").tofile(str(raw))
    f8 = [(f >> 2).astype(np.uint8) for f in fs]                       # 8-bit content, values used as they are
    yuv = tmp_path / "in.yuv"
    with open(yuv, "wb") as fh:
        for f in f8:
            fh.write(f.tobytes())
            fh.write(bytes(256 * 128 // 2))                              # U and V planes
    for path, fmt, src in ((raw, "u16", fs), (yuv, "yuv420p", [f.astype(np.uint16) for f in f8])):
        dec = tmp_path / f"dec_{fmt}.csv"
        r = _run(mip, "-f", "3", "-s", "256x128", "-o", str(path), f"--InputFormat={fmt}", "--NoLog", f"--DecisionsLog={dec}",
                 "--UseAlternativeSamples=1", "--FilterType=filterFrame_1d_int_5x5", "--KernelIdx=2")
        assert r.returncode == 0, r.stdout + r.stderr
        lines = open(dec).read().splitlines()
        assert lines[0] == "POC,CTU,cuSizeName,W,H,CU,X,Y,BestMode,BestCost" and len(lines) - 1 == 3 * 2 * 5380
        for poc in range(3):
            bm, bc = oracle.decisions(oracle.run_frame(src[poc], 5, 2))
            rows = lines[1 + poc * 10760: 1 + (poc + 1) * 10760]
            got_m = np.array([int(x.split(",")[-2]) for x in rows]).reshape(2, 5380)
            got_c = np.array([int(x.split(",")[-1]) for x in rows]).reshape(2, 5380)
            assert np.array_equal(got_m, bm) and np.array_equal(got_c, bc)
            assert rows[0].startswith(f"{poc},0,ALL_AL_64x64,64,64,0,0,0,") and rows[5380].startswith(f"{poc},1,ALL_AL_64x64,64,64,0,128,0,")


@pytest.mark.gpu
def test_cli_topk_energy_and_stage_stamps(mip, oracle, tmp_path):
    """--TopK adds ModeN,CostN columns in (cost, mode) order; --Energy prints joules; the stage stamps that
    computeEnergy_NVIDIA.py:44-96 parses are all present by default and absent with --StageStamps=0."""
    from mipb200 import frames
    fs = [frames.natural_frame(256, 128, 90 + i) for i in range(2)]
    raw = tmp_path / "in.u16"
    np.stack(fs).astype("<u2").tofile(str(raw))
    dec = tmp_path / "dec.csv"
    r = _run(mip, "-f", "2", "-s", "256x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", f"--DecisionsLog={dec}", "--TopK=3", "--Energy")
    assert r.returncode == 0, r.stdout + r.stderr
    lines = open(dec).read().splitlines()
    assert lines[0] == "POC,CTU,cuSizeName,W,H,CU,X,Y,BestMode,BestCost,Mode2,Cost2,Mode3,Cost3" and len(lines) - 1 == 2 * 2 * 5380
    for poc in range(2):
        tm, tc = oracle.topk(oracle.run_frame(fs[poc]), 3)
        rows = np.array([[int(v) for v in x.split(",")[-6:]] for x in lines[1 + poc * 10760: 1 + (poc + 1) * 10760]]).reshape(2, 5380, 3, 2)
        assert np.array_equal(rows[..., 0], tm) and np.array_equal(rows[..., 1], tc)
    for stamp in ("STARTED HOST", "START WRITE SAMPLES MEMOBJ", "FINISH WRITE SAMPLES MEMOBJ", "START ENQUEUE initBoundaries",
                  "FINISH ENQUEUE initBoundaries", "START ENQUEUE reducedPred", "FINISH ENQUEUE reducedPred",
                  "START ENQUEUE upsamplePred_SIZEID=2", "FINISH ENQUEUE upsamplePred_SIZEID=2", "START ENQUEUE upsamplePred_SIZEID=1",
                  "FINISH ENQUEUE upsamplePred_SIZEID=1", "START ENQUEUE upsamplePred_SIZEID=0", "FINISH ENQUEUE upsamplePred_SIZEID=0",
                  "START READ DISTORTION", "FINISH READ DISTORTION"):
        hits = [ln for ln in r.stdout.splitlines() if ln.startswith(stamp + " @ ")]
        assert hits, stamp
        import re
        assert re.fullmatch(r".* @ \d\d:\d\d:\d\d\.\d\d\d", hits[-1]), hits[-1]      # "%H:%M:%S.%f" of the energy script
    assert ("Energy per frame (J)," in r.stdout) or ("Energy counter unavailable" in r.stdout)
    r = _run(mip, "-f", "2", "-s", "256x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", f"--DecisionsLog={dec}", "--StageStamps=0")
    assert r.returncode == 0 and "START ENQUEUE" not in r.stdout and "STARTED HOST" in r.stdout


def test_topk_argument_checks(mip, tmp_path):
    r = _run(mip, "-f", "1", "-s", "128x128", "-o", "x.u16", "--TopK=3")
    assert r.returncode == 1 and "TopK needs --DecisionsLog" in r.stdout
    r = _run(mip, "-f", "1", "-s", "128x128", "-o", "x.u16", "--TopK=40", "--DecisionsLog=d.csv")
    assert r.returncode == 1 and "TopK must be in 1..12" in r.stdout


@pytest.mark.gpu
def test_cli_binary_log_and_threaded_text_log(mip, oracle, tmp_path):
    """--BinaryLog holds every frame's raw table (frames sharded over two workers land at their POC offset); the text log,
    formatted CTU-parallel, is byte-identical to a single pass over the same table."""
    from mipb200 import frames, tables as T
    W, H, N = 640, 384, 3                                    # 15 CTUs: more than one round of formatter threads on small boxes
    fs = [frames.natural_frame(W, H, 300 + i) for i in range(N)]
    raw = tmp_path / "in.u16"
    np.stack(fs).astype("<u2").tofile(str(raw))
    dump = tmp_path / "costs.bin"
    pre = tmp_path / "log"
    import torch
    g = 2 if torch.cuda.device_count() >= 2 else 1
    r = _run(mip, "-f", str(N), "-s", f"{W}x{H}", "-o", str(raw), "--InputFormat=u16", "-l", str(pre), f"--BinaryLog={dump}", f"--NumGpus={g}",
             "--UseAlternativeSamples=1", "--FilterType=filterFrame_2d_int_quarterCtu", "--KernelIdx=3", "--Compat")
    assert r.returncode == 0, r.stdout + r.stderr
    hdr, costs = frames.read_cost_dump(str(dump))
    assert hdr == dict(version=1, width=W, height=H, frames=N, n_ctus=15, costs_per_ctu=97840, bit_depth=10, filter_type=3, kernel_idx=3)
    for poc in range(N):
        assert np.array_equal(costs[poc], oracle.run_frame(fs[poc], 3, 3)), poc
    # text log of frame 0: header + 15 x 97840 lines, in CTU / type / CU / mode order with the right cost in the last column
    lines = open(str(pre) + ".csv").read().splitlines()
    assert lines[0] == "CTU,cuSizeName,W,H,CU,X,Y,Mode,SAD,SATD,minSadHad" and len(lines) - 1 == 15 * 97840
    last = np.array([int(x[x.rfind(",") + 1:]) for x in lines[1:]], dtype=np.int32).reshape(15, 97840)
    assert np.array_equal(last, costs[0])
    ctu_col = np.array([int(x[:x.find(",")]) for x in lines[1::97840]])
    assert ctu_col.tolist() == list(range(15))
    assert lines[1 + 14 * 97840].startswith("14,ALL_AL_64x64,64,64,0,512,256,0,0,0,")


@pytest.mark.gpu
def test_cli_bit_depth(mip, oracle, tmp_path):
    """--BitDepth reaches the engine: a 12-bit frame's table equals the 12-bit oracle (and not the 10-bit one); bad values are refused."""
    from mipb200 import frames
    f = frames.noise_frame(256, 128, 77, bits=12)
    raw = tmp_path / "in.u16"
    f.astype("<u2").tofile(str(raw))
    dump = tmp_path / "c.bin"
    r = _run(mip, "-f", "1", "-s", "256x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", f"--BinaryLog={dump}", "--BitDepth=12")
    assert r.returncode == 0, r.stdout + r.stderr
    hdr, costs = frames.read_cost_dump(str(dump))
    assert hdr["bit_depth"] == 12
    assert np.array_equal(costs[0], oracle.run_frame(f, bit_depth=12)) and not np.array_equal(costs[0], oracle.run_frame(f, bit_depth=10))
    r = _run(mip, "-f", "1", "-s", "256x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", "--BitDepth=9")
    assert r.returncode == 1 and "BitDepth must be 8, 10 or 12" in r.stdout


def test_no_gpu_message_matches_the_reference(mip, tmp_path):
    """On a box without a GPU the CLI says what the reference says when the device index is out of range (main.cpp:225-227)
    and exits 0 like it; with a GPU the banner 'COMPUTING ON GPU i' appears instead (checked in the GPU tests)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    raw = tmp_path / "in.u16"
    np.zeros((128, 128), dtype="<u2").tofile(str(raw))
    r = _run(mip, "-f", "1", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog")
    assert r.returncode == 0 and "Incorrect GPU index. Only 0 GPUs are detected" in r.stdout


@pytest.mark.gpu
def test_gpu_selection_banner(mip, tmp_path):
    raw = tmp_path / "in.u16"
    np.zeros((128, 128), dtype="<u2").tofile(str(raw))
    r = _run(mip, "-f", "1", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog")
    assert r.returncode == 0 and "COMPUTING ON GPU 0" in r.stdout
    r = _run(mip, "-f", "1", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", "--DeviceIndex=64")
    assert r.returncode == 0 and "Incorrect GPU index. Only" in r.stdout and "TIMING RESULTS" not in r.stdout


def test_log_writers_selftest(mip, tmp_path):
    """tests/cli_writers_selftest.cpp: the CLI's buffered, thread-parallel formatters (cost log with and without POC / SAD /
    SATD, decisions log with a top-3 shortlist) write byte for byte what plain fprintf loops write.  No GPU needed."""
    libdir = os.path.join(ROOT, "vvc-mip-gpu_b200", "lib")
    exe = tmp_path / "selftest"
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-o", str(exe), os.path.join(ROOT, "tests", "cli_writers_selftest.cpp"),
                    "-L", libdir, "-lmipb200", f"-Wl,-rpath,{libdir}"], check=True)
    r = subprocess.run([str(exe), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "selftest: ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_cli_streamed_ring_cycled_input_and_digest(mip, oracle, tmp_path):
    """The streaming host on hardware: 40 frames from a 9-frame file through a 4-slot (8-slot with two GPUs) page-locked ring
    filled by the reader thread; raw decisions at their POC offset equal the oracle; the digest repeats with the input's
    period and does not depend on the number of GPUs."""
    import torch
    from mipb200 import frames
    W, H, N, P = 384, 200, 40, 9
    fs = [frames.natural_frame(W, H, 170 + i) if i % 2 else frames.noise_frame(W, H, 170 + i) for i in range(P)]
    raw = tmp_path / "pool.u16"
    np.stack(fs).astype("<u2").tofile(str(raw))
    want = [oracle.decisions(oracle.run_frame(f, 8, 2)) for f in fs]
    F = ["--UseAlternativeSamples=1", "--FilterType=filterFrame_2d_float_5x5_quarterCtu", "--KernelIdx=2"]
    digests = {}
    for g in sorted({1, min(2, torch.cuda.device_count())}):
        dec, dig = tmp_path / f"dec{g}.bin", tmp_path / f"dig{g}.csv"
        r = _run(mip, "-f", str(N), "-s", f"{W}x{H}", "-o", str(raw), "--InputFormat=u16", f"--InputFrames={P}", "--RingFrames=4", "--NoLog",
                 f"--DecisionsBin={dec}", f"--Digest={dig}", f"--NumGpus={g}", "--StageStamps=0", *F)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "streamed by a reader thread" in r.stdout and "page-locked" in r.stdout
        hdr, modes, costs = frames.read_decisions_dump(str(dec))
        assert hdr["frames"] == N and hdr["k"] == 1 and hdr["filter_type"] == 8
        for poc in range(N):
            bm, bc = want[poc % P]
            assert np.array_equal(modes[poc, :, :, 0], bm) and np.array_equal(costs[poc, :, :, 0], bc), (g, poc)
        lines = open(dig).read().splitlines()
        assert lines[0] == "POC,Modes,BestCosts" and len(lines) == N + 1
        for poc in range(P, N):
            assert lines[1 + poc].split(",")[1:] == lines[1 + poc - P].split(",")[1:]
        digests[g] = lines
    assert len({tuple(v) for v in digests.values()}) == 1
    # full tables: --BinaryLog + --Digest carries a Costs column; resident ring (default capacity)
    dump, dig = tmp_path / "c.bin", tmp_path / "digc.csv"
    r = _run(mip, "-f", "12", "-s", f"{W}x{H}", "-o", str(raw), "--InputFormat=u16", f"--InputFrames={P}", "--NoLog", f"--BinaryLog={dump}", f"--Digest={dig}", "--StageStamps=0", *F)
    assert r.returncode == 0 and "resident" in r.stdout, r.stdout + r.stderr
    _, costs = frames.read_cost_dump(str(dump))
    for poc in (0, 5, 11):
        assert np.array_equal(costs[poc], oracle.run_frame(fs[poc % P], 8, 2)), poc
    lines = open(dig).read().splitlines()
    assert lines[0] == "POC,Costs,Modes,BestCosts" and lines[1].split(",")[1:] == lines[10].split(",")[1:]
    assert lines[1].split(",")[2:] == digests[1][1].split(",")[1:]          # same decisions hash as the decisions-only run


@pytest.mark.gpu
def test_cli_rejects_samples_beyond_the_bit_depth(mip, tmp_path):
    from mipb200 import frames
    f = frames.noise_frame(128, 128, 3, bits=12)
    raw = tmp_path / "in.u16"
    f.astype("<u2").tofile(str(raw))
    r = _run(mip, "-f", "1", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog")
    assert r.returncode == 1 and "does not fit 10 bits" in r.stderr


@pytest.mark.gpu
def test_cli_compact_log(mip, oracle, tmp_path):
    """--CompactLog on the GPU: every frame's compact table at its POC offset; expanded it equals the oracle's int32 table."""
    from mipb200 import frames
    W, H, N = 384, 200, 4
    fs = [frames.noise_frame(W, H, 700 + i) for i in range(N)]
    raw = tmp_path / "in.u16"
    np.stack(fs).astype("<u2").tofile(str(raw))
    dump = tmp_path / "c.cmp"
    r = _run(mip, "-f", str(N), "-s", f"{W}x{H}", "-o", str(raw), "--InputFormat=u16", "--NoLog", f"--CompactLog={dump}", "--StageStamps=0",
             "--UseAlternativeSamples=1", "--FilterType=filterFrame_1d_float", "--KernelIdx=4")
    assert r.returncode == 0, r.stdout + r.stderr
    hdr, rec = frames.read_compact_dump(str(dump))
    assert hdr["frames"] == N and hdr["bytes_per_ctu"] == mip.COMPACT_BYTES_PER_CTU and hdr["filter_type"] == 2
    for poc in range(N):
        assert np.array_equal(mip.expand_costs(rec[poc]), oracle.run_frame(fs[poc], 2, 4)), poc
