"""GPU: the C++ drivers (reference class interface + file formats on top of libictrack.so) against golden outputs
of the reference's OWN main()s (tests/golden/drivers.npz, made by oracle/_ref/run_track_nposes and
oracle/_ref/run_io_reprojection_test, see tests/golden/make_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "invcompcamtrack_b200", "bin")
GOLD = np.load(os.path.join(ROOT, "tests", "golden", "drivers.npz"))


def write_pgm(path, img):
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (img.shape[1], img.shape[0]))
        f.write(np.ascontiguousarray(img, np.uint8).tobytes())


@pytest.fixture(scope="module")
def drivers():
    from invcompcamtrack_b200 import build
    build.build_library()
    exes = build.build_drivers()
    assert len(exes) == 2
    return BIN


def parse(txt):
    return [[float(v) for v in line.split()] for line in txt.strip().split("\n")]


def test_run_track_nposes_matches_reference_driver(drivers, tmp_path):
    d = str(tmp_path)
    for k, f in enumerate(GOLD["frames"]):
        write_pgm(os.path.join(d, "f%d.pgm" % k), f)
    open(os.path.join(d, "in.txt"), "w").write(str(GOLD["nposes_input"]).replace("@DIR@", d))
    subprocess.run([os.path.join(drivers, "run_track_nposes"), os.path.join(d, "in.txt"), os.path.join(d, "out.txt")],
                   check=True)                                    # reference summation order is the library default
    got, ref = open(os.path.join(d, "out.txt")).read(), str(GOLD["nposes_output"])
    g, r = parse(got), parse(ref)
    assert [len(x) for x in g] == [len(x) for x in r]
    for lg, lr in zip(g, r):
        if len(lr) == 6:                     # a pose line, precision 8: reference order => same digits
            assert lg == lr
        else:                                # an NCC line, precision 3 (the GPU sums the patches as a tree)
            assert np.abs(np.array(lg) - np.array(lr)).max() <= 2e-3
    # fast mode (tree order, ICT_SUM_ORDER=0): same poses up to fp32 summation noise
    subprocess.run([os.path.join(drivers, "run_track_nposes"), os.path.join(d, "in.txt"), os.path.join(d, "out2.txt")],
                   check=True, env=dict(os.environ, ICT_SUM_ORDER="0"))
    for lg, lr in zip(parse(open(os.path.join(d, "out2.txt")).read()), r):
        assert np.abs(np.array(lg) - np.array(lr)).max() <= (1e-4 if len(lr) == 6 else 2e-3)


def test_run_track_nposes_state_between_frame_steps(drivers, tmp_path):
    """ResetOdometer runs from Set3Dpoints only (odometer.cpp:153, 173), so a point that leaves the image
    mid-chain keeps its last template patch and steepest-descent values — also from an earlier frame step — and they
    keep feeding the Hessian.  Fixture: a pan in which a quarter of the correspondences cross the left / right border
    (tests/golden/drivers_stale.npz, made by the reference's own run_track_nposes main).  The GPU driver (tracker knob
    "keep_state", on by default in the driver) must print the same pose digits; without the carried state it does not."""
    gold = np.load(os.path.join(ROOT, "tests", "golden", "drivers_stale.npz"))
    d = str(tmp_path)
    for k, f in enumerate(gold["frames"]):
        write_pgm(os.path.join(d, "f%d.pgm" % k), f)
    open(os.path.join(d, "in.txt"), "w").write(str(gold["nposes_input"]).replace("@DIR@", d))
    subprocess.run([os.path.join(drivers, "run_track_nposes"), os.path.join(d, "in.txt"), os.path.join(d, "out.txt")], check=True)
    g, r = parse(open(os.path.join(d, "out.txt")).read()), parse(str(gold["nposes_output"]))
    assert [len(x) for x in g] == [len(x) for x in r]
    for lg, lr in zip(g, r):
        if len(lr) == 6:
            assert lg == lr                  # pose lines, precision 8: digit for digit
        else:
            assert np.abs(np.array(lg) - np.array(lr)).max() <= 2e-3
    # and the state matters on this fixture: with every frame step starting from zeroed arrays some pose differs
    env = dict(os.environ, ICT_KEEP_STATE="0")
    subprocess.run([os.path.join(drivers, "run_track_nposes"), os.path.join(d, "in.txt"), os.path.join(d, "out0.txt")],
                   check=True, env=env)
    g0 = parse(open(os.path.join(d, "out0.txt")).read())
    assert any(lg != lr for lg, lr in zip(g0, r) if len(lr) == 6)


def test_run_track_matches_reference_driver(drivers, tmp_path):
    d = str(tmp_path)
    write_pgm(os.path.join(d, "a.pgm"), GOLD["frames"][0])
    write_pgm(os.path.join(d, "b.pgm"), GOLD["frames"][1])
    open(os.path.join(d, "pair.bin"), "wb").write(GOLD["pair_input"].tobytes())
    args = str(GOLD["pair_args"]).split()
    subprocess.run([os.path.join(drivers, "run_track"), os.path.join(d, "a.pgm"), os.path.join(d, "b.pgm"),
                    os.path.join(d, "pair.bin"), os.path.join(d, "pair.out")] + args, check=True)
    got = np.frombuffer(open(os.path.join(d, "pair.out"), "rb").read(), "<f8")
    assert np.array_equal(got, GOLD["pair_output"])
    subprocess.run([os.path.join(drivers, "run_track"), os.path.join(d, "a.pgm"), os.path.join(d, "b.pgm"),
                    os.path.join(d, "pair.bin"), os.path.join(d, "pair2.out")] + args, check=True,
                   env=dict(os.environ, ICT_SUM_ORDER="0"))
    got2 = np.frombuffer(open(os.path.join(d, "pair2.out"), "rb").read(), "<f8")
    assert np.abs(got2 - GOLD["pair_output"]).max() < 1e-5
