"""The drop-in proof: the reference's own cqb3.cpp, unmodified, compiled against this repository's
header and linked to its libQB3.so (tests/dropin/Makefile; libicd replaced by a PNM stand-in), converts files both
ways, and so does the same source compiled against the REFERENCE's QB3.h (QB3_MAXBANDS 16) -- the binary a user
already has. The streams it writes are the oracle's, byte for byte."""
import os
import subprocess

import numpy as np
import pytest

from helpers import MODE_BASE, MODE_BEST, MODE_FTL, content, oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "tests", "dropin")
BINARIES = ("cqb3_new_header", "cqb3_ref_header")


def test_reference_cli_builds_against_this_library():
    """CPU side: the reference's translation unit compiles and links (needs /root/reference, so only where it exists)."""
    if not os.path.exists("/root/reference/cqb3.cpp"):
        pytest.skip("the reference sources are not on this box; the binaries were built where they are")
    subprocess.run(["make", "-s", "-C", DROPIN], check=True)
    for b in BINARIES:
        path = os.path.join(DROPIN, "_build", b)
        assert os.path.exists(path)
        undefined = subprocess.run(["nm", "-D", "--undefined-only", path], capture_output=True, text=True).stdout
        used = sorted({l.split()[-1] for l in undefined.splitlines() if " qb3_" in l})
        assert "qb3_encode" in used and "qb3_read_data" in used and len(used) >= 14, used


def write_pnm(path, img):
    h, w, b = img.shape
    with open(path, "wb") as f:
        f.write(b"P%d\n%d %d\n%d\n" % (6 if b == 3 else 5, w, h, 255 if img.dtype == np.uint8 else 65535))
        f.write(img.astype(">u2").tobytes() if img.dtype == np.uint16 else img.tobytes())


def read_pnm(path, dtype):
    data = open(path, "rb").read()
    parts = data.split(b"\n", 3)
    w, h = (int(v) for v in parts[1].split())
    b = 3 if parts[0] == b"P6" else 1
    pix = np.frombuffer(parts[3], dtype=">u2" if dtype == np.uint16 else np.uint8)
    return pix.astype(dtype).reshape(h, w, b)


@pytest.mark.gpu
@pytest.mark.parametrize("binary", BINARIES)
def test_reference_cli_round_trips_on_this_library(binary, tmp_path):
    exe = os.path.join(DROPIN, "_build", binary)
    if not os.path.exists(exe):
        pytest.skip("tests/dropin/_build is not built (make -C tests/dropin where /root/reference exists)")
    for (w, h, b, dt, flags, mode) in [(64, 48, 3, np.uint8, [], MODE_BASE), (64, 48, 3, np.uint8, ["-f"], MODE_FTL),
                                       (37, 21, 1, np.uint8, ["-b"], MODE_BEST), (40, 32, 3, np.uint16, [], MODE_BASE)]:
        img = content("synth", w, h, b, dt)
        src, qb3, back = (str(tmp_path / n) for n in ("in.pnm", "out.qb3", "back.pnm"))
        write_pnm(src, img)
        subprocess.run([exe] + flags + [src, qb3], check=True, capture_output=True)
        stream = open(qb3, "rb").read()
        assert stream == oracle().encode(img, mode=mode), (binary, w, h, b, dt, flags)
        subprocess.run([exe, "-d", qb3, back], check=True, capture_output=True)
        assert np.array_equal(read_pnm(back, dt), img)
