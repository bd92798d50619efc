"""INTEGRATION Route B (integration/route_b/optical_flow_klt_b200.h): GPU subclasses of the REFERENCE's own tracker classes that
override the private virtuals TrackMultipleLevel / TrackSingleLevel (optical_flow.h:83-86).  The header is compiled against the
reference headers where they lie (+ oracle/shim for the absent Slam_Utility / Eigen) and linked with libftk_b200.so by
`make -C oracle route_b` -> oracle/_ref/route_b_test (a built artefact: it travels to the GPU box, the reference tree does not).

CPU: the recipe builds.  GPU: for every variant x method, multi- and single-level, the B200 subclass returns bit-identical positions
and identical status to the reference's CPU class running in the same process through the same public TrackFeatures."""
import os
import stat
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "oracle", "_ref", "route_b_test")
HAVE_REFERENCE = os.path.isdir("/root/reference")


@pytest.mark.skipif(not HAVE_REFERENCE, reason="needs the reference headers (only in the build container)")
def test_route_b_builds_against_reference_headers():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "route_b"])
    assert os.path.exists(EXE)
    text = open(os.path.join(ROOT, "integration", "route_b", "optical_flow_klt_b200.h")).read()
    for name in ("TrackMultipleLevel", "TrackSingleLevel", "override", "OpticalFlowBasicKltB200", "OpticalFlowAffineKltB200", "OpticalFlowLssdKltB200"):
        assert name in text


@pytest.mark.gpu
def test_route_b_matches_reference_classes(tmp_path):
    if not os.path.exists(EXE):
        if not HAVE_REFERENCE:
            pytest.skip("oracle/_ref/route_b_test was not built (needs the reference headers)")
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "route_b"])
    os.chmod(EXE, os.stat(EXE).st_mode | stat.S_IXUSR)
    from feature_tracker_b200 import synthetic as S
    rows, cols, levels, n = 240, 320, 4, 160
    ref, cur, uv, _ = S.make_pair(rows, cols, n, pair_id=77, border=10)
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(struct.pack("4i", rows, cols, levels, n))
        for a in (ref, cur, uv):
            f.write(np.ascontiguousarray(a).tobytes())
    subprocess.check_call([EXE, fin, fout], timeout=300)
    data = open(fout, "rb").read()
    off = 0

    def take(dtype, count):
        nonlocal off
        a = np.frombuffer(data, dtype=dtype, count=count, offset=off)
        off += a.nbytes
        return a

    tracked_total = 0
    for variant in ("basic", "affine", "lssd"):
        for method in ("inverse", "direct", "fast"):
            for mode in ("multi", "single"):
                ok_cpu, ok_gpu = take(np.int32, 2)
                cpu_uv, cpu_st = take(np.uint32, 2 * n), take(np.uint8, n)
                gpu_uv, gpu_st = take(np.uint32, 2 * n), take(np.uint8, n)
                tag = f"{variant}/{method}/{mode}"
                assert ok_cpu == 1 and ok_gpu == 1, tag
                assert (cpu_st == gpu_st).all(), tag
                assert (cpu_uv == gpu_uv).all(), tag
                tracked_total += int((gpu_st == 1).sum())
    assert off == len(data) and tracked_total > 9 * n  # the comparison is not vacuous
