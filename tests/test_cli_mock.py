"""The CLI end to end WITHOUT a GPU: csrc/main.cpp linked against tests/mock_engine/mock_mipb200.cpp, a test double of the C
ABI that answers with the CPU oracle.  What is tested here is the host plumbing of the CLI only -- options, worker threads
and frame sharding, result handling, the text / decisions / binary log writers, the stamps -- never the CUDA engine (that is
what the `-m gpu` tests do with the real library).  The mock is built into a temporary directory and never shipped."""
import os
import re
import subprocess

import numpy as np
import pytest

from mipb200 import frames

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def mock_cli(tmp_path_factory, oracle):
    d = tmp_path_factory.mktemp("mockcli")
    exe = d / "mipb200_main_mock"
    odir = os.path.join(ROOT, "oracle")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", "-pthread", "-fopenmp", "-o", str(exe),
                    os.path.join(ROOT, "vvc-mip-gpu_b200", "csrc", "main.cpp"), os.path.join(ROOT, "tests", "mock_engine", "mock_mipb200.cpp"),
                    "-L", odir, "-lmip_oracle", f"-Wl,-rpath,{odir}"], check=True)

    def run(*args, gpus=1):
        env = dict(os.environ, MOCK_GPUS=str(gpus), OMP_NUM_THREADS="2")
        return subprocess.run([str(exe), *args], capture_output=True, text=True, timeout=900, env=env)

    return run


def test_cost_log_of_frame_0(mock_cli, oracle, tmp_path):
    fs = [frames.natural_frame(256, 136, 40 + i) for i in range(2)]
    csv = tmp_path / "in.csv"
    frames.write_csv(str(csv), fs)
    r = mock_cli("-f", "2", "-s", "256x136", "-o", str(csv), "-l", str(tmp_path / "log"), "--UseAlternativeSamples=1",
                 "--Filter=filterFrame_1d_float", "--KernelIdx=3")
    assert r.returncode == 0, r.stdout + r.stderr
    assert "COMPUTING ON GPU 0" in r.stdout and "Current frame 1" in r.stdout and "TIMING RESULTS (miliseconds)" in r.stdout
    cost, sad, satd = oracle.run_frame(fs[0], 2, 3, want_sad_satd=True)
    lines = open(str(tmp_path / "log") + ".csv").read().splitlines()
    assert lines[0] == "CTU,cuSizeName,W,H,CU,X,Y,Mode,SAD,SATD,minSadHad" and len(lines) - 1 == 4 * 97840
    got = np.array([[int(v) for v in x.rsplit(",", 3)[1:]] for x in lines[1:]], dtype=np.int64).reshape(4, 97840, 3)
    assert np.array_equal(got[..., 0], sad) and np.array_equal(got[..., 1], satd) and np.array_equal(got[..., 2], cost)
    assert lines[1].startswith("0,ALL_AL_64x64,64,64,0,0,0,0,") and lines[1 + 3 * 97840].startswith("3,ALL_AL_64x64,64,64,0,128,128,0,")
    for stamp in ("STARTED HOST", "START BUILD KERNELS", "FINISH BUILD KERNELS", "START WRITE SAMPLES MEMOBJ", "START ENQUEUE filterFrame",
                  "START ENQUEUE upsamplePred_SIZEID=0", "FINISH READ DISTORTION", "FINISHED HOST"):
        assert re.search(r"^" + re.escape(stamp) + r" @ \d\d:\d\d:\d\d\.\d\d\d$", r.stdout, re.M), stamp


def test_all_frames_two_workers_binary_and_decision_logs(mock_cli, oracle, tmp_path):
    """Three frames sharded over two (mock) GPUs: the POC-prefixed text log, the raw dump and the top-3 decisions log all
    come out in POC order and equal the oracle."""
    W, H, N = 256, 128, 3
    fs = [frames.noise_frame(W, H, 50 + i) for i in range(N)]
    raw = tmp_path / "in.u16"
    np.stack(fs).astype("<u2").tofile(str(raw))
    pre, dump, dec = tmp_path / "all", tmp_path / "c.bin", tmp_path / "dec.csv"
    r = mock_cli("-f", str(N), "-s", f"{W}x{H}", "-o", str(raw), "--InputFormat=u16", "-l", str(pre), "--AllFrames", "--Compat", "--NumGpus=2",
                 f"--BinaryLog={dump}", f"--DecisionsLog={dec}", "--TopK=3", gpus=2)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "COMPUTING ON GPU 0" in r.stdout and "COMPUTING ON GPU 1" in r.stdout and "on 2 GPU(s)" in r.stdout
    want = [oracle.run_frame(f) for f in fs]
    hdr, costs = frames.read_cost_dump(str(dump))
    assert hdr["frames"] == N and hdr["n_ctus"] == 2 and hdr["bit_depth"] == 10
    for poc in range(N):
        assert np.array_equal(costs[poc], want[poc])
    lines = open(str(pre) + ".csv").read().splitlines()
    assert lines[0] == "POC,CTU,cuSizeName,W,H,CU,X,Y,Mode,SAD,SATD,minSadHad" and len(lines) - 1 == N * 2 * 97840
    tail = np.array([[int(v) for v in x.rsplit(",", 3)[1:]] for x in lines[1:]], dtype=np.int64).reshape(N, 2, 97840, 3)
    assert not tail[..., :2].any() and np.array_equal(tail[..., 2], np.stack(want))          # --Compat: SAD / SATD columns print 0
    assert [int(x.split(",", 1)[0]) for x in lines[1::2 * 97840]] == list(range(N))
    dl = open(dec).read().splitlines()
    assert dl[0] == "POC,CTU,cuSizeName,W,H,CU,X,Y,BestMode,BestCost,Mode2,Cost2,Mode3,Cost3" and len(dl) - 1 == N * 2 * 5380
    for poc in range(N):
        tm, tc = oracle.topk(want[poc], 3)
        rows = np.array([[int(v) for v in x.split(",")[-6:]] for x in dl[1 + poc * 10760: 1 + (poc + 1) * 10760]]).reshape(2, 5380, 3, 2)
        assert np.array_equal(rows[..., 0], tm) and np.array_equal(rows[..., 1], tc)
        assert dl[1 + poc * 10760].startswith(f"{poc},0,ALL_AL_64x64,64,64,0,0,0,")


def test_device_index_out_of_range_and_bit_depth(mock_cli, oracle, tmp_path):
    f = frames.noise_frame(128, 128, 9, bits=12)
    raw = tmp_path / "in.u16"
    f.astype("<u2").tofile(str(raw))
    r = mock_cli("-f", "1", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", "--BitDepth=12", gpus=0)
    assert r.returncode == 0 and "Incorrect GPU index. Only 0 GPUs are detected" in r.stdout and "TIMING RESULTS" not in r.stdout
    r = mock_cli("-f", "1", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", "--BitDepth=12", "--NumGpus=3", gpus=2)
    assert r.returncode == 0 and "Incorrect GPU index. Only 2 GPUs are detected" in r.stdout
    dump = tmp_path / "c.bin"
    r = mock_cli("-f", "1", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", f"--BinaryLog={dump}", "--BitDepth=12", "--Energy")
    assert r.returncode == 0 and "Energy counter unavailable" in r.stdout
    hdr, costs = frames.read_cost_dump(str(dump))
    assert hdr["bit_depth"] == 12 and np.array_equal(costs[0], oracle.run_frame(f, bit_depth=12))


def test_input_formats(mock_cli, oracle, tmp_path):
    """csv, raw u16, 8-bit yuv420p and 16-bit yuv420p10le inputs of the same two frames give the same table (chroma planes
    are skipped; 8-bit samples are taken as they are)."""
    W, H = 136, 72
    f10 = [frames.natural_frame(W, H, 60 + i) for i in range(2)]
    f8 = [(f >> 2).astype(np.uint16) for f in f10]
    files = {}
    p = tmp_path / "a.csv"; frames.write_csv(str(p), f10); files["csv"] = (p, f10)
    p = tmp_path / "a.u16"; np.stack(f10).astype("<u2").tofile(str(p)); files["u16"] = (p, f10)
    p = tmp_path / "a10.yuv"
    with open(p, "wb") as fh:
        for f in f10:
            fh.write(f.astype("<u2").tobytes()); fh.write(bytes(2 * (W * H // 2)))      # Y, then U and V (16-bit, quarter size each)
    files["yuv420p10le"] = (p, f10)
    p = tmp_path / "a8.yuv"
    with open(p, "wb") as fh:
        for f in f8:
            fh.write(f.astype(np.uint8).tobytes()); fh.write(bytes(W * H // 2))
    files["yuv420p"] = (p, f8)
    for fmt, (path, src) in files.items():
        dump = tmp_path / f"{fmt}.bin"
        r = mock_cli("-f", "2", "-s", f"{W}x{H}", "-o", str(path), f"--InputFormat={fmt}", "--NoLog", f"--BinaryLog={dump}", "--StageStamps=0")
        assert r.returncode == 0, fmt + r.stdout + r.stderr
        assert "START ENQUEUE" not in r.stdout
        _, costs = frames.read_cost_dump(str(dump))
        for poc in range(2):
            assert np.array_equal(costs[poc], oracle.run_frame(src[poc])), (fmt, poc)
    r = mock_cli("-f", "3", "-s", f"{W}x{H}", "-o", str(files["u16"][0]), "--InputFormat=u16", "--NoLog")
    assert r.returncode == 1 and "holds fewer than 3 frames" in r.stderr


def test_samples_must_fit_the_bit_depth(mock_cli, tmp_path):
    """A sample >= 1 << BitDepth would overflow the engine's packed 16-bit arithmetic silently: the readers refuse it and say where it is."""
    f = frames.noise_frame(128, 128, 9, bits=10)
    f[5, 77] = 1024
    raw = tmp_path / "in.u16"
    f.astype("<u2").tofile(str(raw))
    r = mock_cli("-f", "1", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog")
    assert r.returncode == 1 and "sample 1024 at row 5, column 77 does not fit 10 bits" in r.stderr
    csv = tmp_path / "in.csv"
    frames.write_csv(str(csv), [f])
    r = mock_cli("-f", "1", "-s", "128x128", "-o", str(csv), "--NoLog")
    assert r.returncode == 1 and "sample 1024 at row 5, column 77 does not fit 10 bits" in r.stderr
    r = mock_cli("-f", "1", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", "--BitDepth=12")
    assert r.returncode == 0, r.stdout + r.stderr


def test_streamed_ring_cycled_input_binary_decisions_and_digest(mock_cli, oracle, tmp_path):
    """Twenty frames from a nine-frame file (--InputFrames 9) through a four- or eight-slot ring: the reader thread streams and rewinds,
    two workers shard the frames, decisions land raw at their POC offset, and the digest is the same for one and two workers
    and repeats with the input's period."""
    W, H, N, P = 136, 72, 20, 9
    fs = [frames.natural_frame(W, H, 70 + i) for i in range(P)]
    raw = tmp_path / "pool.u16"
    np.stack(fs).astype("<u2").tofile(str(raw))
    want = [oracle.decisions(oracle.run_frame(f, 3, 1)) for f in fs]
    digests = {}
    for g in (1, 2):
        dec, dig = tmp_path / f"dec{g}.bin", tmp_path / f"dig{g}.csv"
        r = mock_cli("-f", str(N), "-s", f"{W}x{H}", "-o", str(raw), "--InputFormat=u16", f"--InputFrames={P}", "--RingFrames=4", "--NoLog",
                     f"--DecisionsBin={dec}", f"--Digest={dig}", f"--NumGpus={g}", "--UseAlternativeSamples=1",
                     "--FilterType=filterFrame_2d_int_quarterCtu", "--KernelIdx=1", gpus=2)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "streamed by a reader thread" in r.stdout and f"Frame ring: {4 * g} x" in r.stdout
        hdr, modes, costs = frames.read_decisions_dump(str(dec))
        assert hdr == dict(version=1, width=W, height=H, frames=N, n_ctus=2, cus_per_ctu=5380, bit_depth=10, filter_type=3, kernel_idx=1, k=1)
        for poc in range(N):
            bm, bc = want[poc % P]
            assert np.array_equal(modes[poc, :, :, 0], bm) and np.array_equal(costs[poc, :, :, 0], bc), (g, poc)
        lines = open(dig).read().splitlines()
        assert lines[0] == "POC,Modes,BestCosts" and [ln.split(",")[0] for ln in lines[1:]] == [str(i) for i in range(N)]
        digests[g] = lines
        for poc in range(P, N):
            assert lines[1 + poc].split(",")[1:] == lines[1 + poc - P].split(",")[1:]
        assert len({ln.split(",", 1)[1] for ln in lines[1:1 + P]}) == P          # distinct frames hash differently
    assert digests[1] == digests[2]
    # a resident ring (the default: the nine distinct frames fit) gives the same bytes
    dec = tmp_path / "dec_res.bin"
    r = mock_cli("-f", str(N), "-s", f"{W}x{H}", "-o", str(raw), "--InputFormat=u16", f"--InputFrames={P}", "--NoLog", f"--DecisionsBin={dec}",
                 "--UseAlternativeSamples=1", "--FilterType=filterFrame_2d_int_quarterCtu", "--KernelIdx=1")
    assert r.returncode == 0 and "resident" in r.stdout
    assert open(dec, "rb").read() == open(tmp_path / "dec1.bin", "rb").read()


def test_streamed_csv_input_and_all_frames_logs(mock_cli, oracle, tmp_path):
    """A CSV larger than the ring is parsed incrementally; the POC-ordered text logs of a streamed, two-worker run equal the oracle."""
    W, H, N = 136, 72, 5
    fs = [frames.noise_frame(W, H, 90 + i) for i in range(N)]
    csv = tmp_path / "in.csv"
    frames.write_csv(str(csv), fs)
    pre, dec = tmp_path / "all", tmp_path / "dec.csv"
    r = mock_cli("-f", str(N), "-s", f"{W}x{H}", "-o", str(csv), "--RingFrames=2", "--NumGpus=2", "-l", str(pre), "--AllFrames", "--Compat",
                 f"--DecisionsLog={dec}", gpus=2)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Frame ring: 5 x" in r.stdout and "resident" in r.stdout          # at least four frames per GPU: all five fit
    r = mock_cli("-f", str(N), "-s", f"{W}x{H}", "-o", str(csv), "--RingFrames=2", "-l", str(pre), "--AllFrames", "--Compat", f"--DecisionsLog={dec}")
    assert r.returncode == 0 and "Frame ring: 4 x" in r.stdout and "streamed" in r.stdout
    want = [oracle.run_frame(f) for f in fs]
    lines = open(str(pre) + ".csv").read().splitlines()
    assert len(lines) - 1 == N * 2 * 97840
    got = np.array([int(x[x.rfind(",") + 1:]) for x in lines[1:]], dtype=np.int32).reshape(N, 2, 97840)
    assert np.array_equal(got, np.stack(want))
    assert [int(x.split(",", 1)[0]) for x in lines[1::2 * 97840]] == list(range(N))
    dl = open(dec).read().splitlines()
    assert len(dl) - 1 == N * 2 * 5380
    for poc in range(N):
        bm, bc = oracle.decisions(want[poc])
        rows = dl[1 + poc * 10760: 1 + (poc + 1) * 10760]
        assert np.array_equal(np.array([int(x.split(",")[-2]) for x in rows]).reshape(2, 5380), bm)
        assert np.array_equal(np.array([int(x.split(",")[-1]) for x in rows]).reshape(2, 5380), bc)
    # a short CSV is still an error, reported with the line count
    r = mock_cli("-f", str(N + 1), "-s", f"{W}x{H}", "-o", str(csv), "--RingFrames=2", "--NoLog")
    assert r.returncode == 1 and f"holds {N * H} lines, need {(N + 1) * H}" in r.stderr


def test_compact_log(mock_cli, oracle, tmp_path, mip):
    """--CompactLog: the compact table of every frame at its POC offset; expanded (mipb200_expand_costs) it is the int32 table."""
    W, H, N = 136, 72, 3
    fs = [frames.natural_frame(W, H, 20 + i) for i in range(N)]
    raw = tmp_path / "in.u16"
    np.stack(fs).astype("<u2").tofile(str(raw))
    dump, dig = tmp_path / "c.cmp", tmp_path / "d.csv"
    r = mock_cli("-f", str(N), "-s", f"{W}x{H}", "-o", str(raw), "--InputFormat=u16", "--NoLog", f"--CompactLog={dump}", f"--Digest={dig}", "--NumGpus=2", gpus=2)
    assert r.returncode == 0, r.stdout + r.stderr
    hdr, rec = frames.read_compact_dump(str(dump))
    assert hdr["frames"] == N and hdr["n_ctus"] == 2 and hdr["bytes_per_ctu"] == mip.COMPACT_BYTES_PER_CTU == 276672
    for poc in range(N):
        assert np.array_equal(mip.expand_costs(rec[poc], threads=2), oracle.run_frame(fs[poc])), poc
    assert open(dig).read().splitlines()[0] == "POC,CostsCompact,Modes,BestCosts"
    r = mock_cli("-f", "1", "-s", f"{W}x{H}", "-o", str(raw), "--InputFormat=u16", f"--CompactLog={dump}")
    assert r.returncode == 1 and "CompactLog needs --NoLog" in r.stdout


def test_slots_and_ring_arguments(mock_cli, tmp_path):
    raw = tmp_path / "in.u16"
    np.zeros((2, 128, 128), dtype="<u2").tofile(str(raw))
    r = mock_cli("-f", "2", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", "--Slots=0")
    assert r.returncode == 1 and "Slots must be in 1..16" in r.stdout
    r = mock_cli("-f", "2", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", "--Slots=5", "--InputFrames=1")
    assert r.returncode == 0 and "Frame ring: 1 x" in r.stdout and "resident" in r.stdout and "Peak host memory (MB)," in r.stdout
    r = mock_cli("-f", "2", "-s", "128x128", "-o", str(raw), "--InputFormat=u16", "--NoLog", "--RingFrames=-1")
    assert r.returncode == 1 and "must be positive" in r.stdout
