"""Host-side proofs behind two device shortcuts (tools/checks/*.c), kept in the CPU suite.

1. The error band of the centred NN filter (kBandCentredA / kBandCentredX, csrc/icpb_internal.h),
hunted for counterexamples on the CPU: tools/checks/filter_band_check.c reproduces the kernels' filter value and the
reference's distance (icp.cpp:606-620) operation for operation.  The long run is profiles/r02_check_filter_band.txt;
this keeps a short one in the suite and ties the program's constants to the header's.
2. The five-instruction FMA division of the back-projection kernel (csrc/cloud.cu, div_by) against `/` for every depth
and every column / row it can meet (pointcloud.cpp:37-39) -- the host twin of the device sweep in test_gpu_cloud.py.
3. The brick-jumping ray walk (csrc/map.cu, map_rays_brick_kernel) as pure integer arithmetic against the plain
voxel-by-voxel walk of the oracle, on random rays, occupancy and z-slabs (tools/checks/brick_walk_emulation.py).
No GPU."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "checks", "filter_band_check.c")
HDR = os.path.join(ROOT, "icp-slam-prototype_b200", "csrc", "icpb_internal.h")


def _build(tmp_path, src=SRC, name="filter_band_check"):
    exe = str(tmp_path / name)
    try:
        has_fma = " fma " in open("/proc/cpuinfo").read()
    except OSError:
        has_fma = False
    # without the FMA instruction fmaf() goes through libm: the same values, slower
    flags = ["-mfma"] if has_fma else []
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", *flags, "-fopenmp", "-o", exe, src, "-lm"])
    return exe


def _run(exe, samples, scale=None):
    cmd = [exe, str(samples)] + ([str(scale)] if scale is not None else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    rows = [ln.split() for ln in r.stdout.splitlines() if ln.startswith("  ")]
    return r.returncode, [(int(c[-4]), int(c[-3]), int(c[-2]), float(c[-1])) for c in rows]


def test_check_uses_the_constants_the_library_ships():
    hdr, src = open(HDR).read(), open(SRC).read()
    for name, short in (("kBandCentredA", "kA"), ("kBandCentredX", "kX")):
        h = re.search(r"%s\s*=\s*([0-9.]+)f\s*\*\s*([0-9.e+-]+)f" % name, hdr)
        s = re.search(r"%s\s*=\s*([0-9.]+)f\s*\*\s*([0-9.e+-]+)f" % short, src)
        assert h and s and h.groups() == s.groups(), (name, h and h.groups(), s and s.groups())
    assert re.search(r"kBandAbs\s*=\s*1\.0e-30f", hdr) and "kAbs = 1.0e-30f" in src
    assert "1.000001f" in src and "1.000001f" in open(os.path.join(os.path.dirname(HDR), "nn.cu")).read()


def test_band_holds_and_the_hunt_has_teeth(tmp_path):
    exe = _build(tmp_path)
    rc, rows = _run(exe, 2000000)
    assert rc == 0 and len(rows) == 7
    assert all(v == 0 for _, v, _, _ in rows), rows                 # nothing declared farther that the reference keeps
    assert sum(d for d, _, _, _ in rows) > 1000000 and all(m > 100000 for _, _, m, _ in rows)
    assert max(t for *_, t in rows) < 0.5, rows                      # the band is not grazed
    # control: a tenth of the band -- pairs the reference ties now fall outside it
    rc, rows = _run(exe, 2000000, 0.1)
    assert max(t for *_, t in rows) > 1.0, rows


def test_fma_division_is_correctly_rounded_for_every_input_of_the_back_projection(tmp_path):
    src = os.path.join(ROOT, "tools", "checks", "fast_div_check.c")
    cu = open(os.path.join(os.path.dirname(HDR), "cloud.cu")).read()
    # the sequence checked is the sequence shipped
    body = re.search(r"float q = __fmul_rn\(a, d\.rc\);(.*?)return __fmaf_rn\(r, d\.rc, q\);", cu, re.S)
    assert body and body.group(1).count("__fmaf_rn") == 3
    exe = _build(tmp_path, src, "fast_div_check")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout[-500:]
    assert int(r.stdout.split()[1]) > 2000000000


def test_brick_jump_walk_reads_the_voxels_the_plain_walk_reads():
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "checks", "brick_walk_emulation.py"), "3000"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "3000 cases, 0 mismatches" in r.stdout, r.stdout[-500:] + r.stderr[-500:]
