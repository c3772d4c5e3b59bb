"""Packaging and command line tool. CPU only: what needs no device -- the consumer program of
cmake/consumer built against the in-tree library, the CMake lists naming every source, cqb3cu -i on golden streams.
The full `cmake --build` + install + find_package round trip runs when QB3_TEST_CMAKE=1 (it recompiles every kernel)."""
import glob
import json
import os
import shutil
import subprocess

import pytest

from helpers import PRODUCT_SO, ROOT, golden_cases


def _tool():
    exe = os.path.join(ROOT, "apps", "cqb3cu")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "apps")], check=True)
    return exe


def test_cmake_lists_name_every_source_and_header():
    text = open(os.path.join(ROOT, "CMakeLists.txt")).read()
    for src in glob.glob(os.path.join(ROOT, "qb3_b200", "csrc", "*.cu")):
        assert "qb3_b200/csrc/" + os.path.basename(src) in text, src
    for hdr in ("include/QB3.h", "include/qb3cu.h"):
        assert hdr in text
    assert 'PREFIX ""' in text and "QB3::libQB3" in text and "100a" in text


def test_consumer_program_links_and_runs(tmp_path):
    """The program a user of the reference would write, compiled against include/ and the in-tree libQB3.so."""
    exe = str(tmp_path / "consumer")
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "cmake", "consumer", "consumer.cpp"),
                    "-o", exe, PRODUCT_SO, "-Wl,-rpath," + os.path.dirname(PRODUCT_SO)], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "QB3 package ok" in out.stdout, out


def test_cli_info_matches_golden_headers(tmp_path):
    exe = _tool()
    names = {0: "uint8", 1: "int8", 2: "uint16", 3: "int16", 4: "uint32", 5: "int32", 6: "uint64", 7: "int64"}
    seen = 0
    for case in golden_cases():
        if case["kind"] != "small" or seen >= 25:
            continue
        p = tmp_path / (case["name"] + ".qb3")
        p.write_bytes(bytes.fromhex(case["stream"]))
        out = subprocess.run([exe, "-i", str(p)], capture_output=True, text=True)
        assert out.returncode == 0, out
        info = json.loads(out.stdout)
        s = bytes.fromhex(case["stream"])
        assert info["xsize"] == 1 + s[4] + 256 * s[5] and info["ysize"] == 1 + s[6] + 256 * s[7] and info["nbands"] == 1 + s[8]
        assert info["dtype"] == names[s[9]]
        assert info["mode"] in ("base_z", "cf", "rle", "cf_rle", "base", "cf_h", "rle_h", "best", "ftl", "stored")
        if case.get("quanta", 1) > 1:
            assert info["quanta"] == case["quanta"]
        seen += 1
    assert seen > 5
    bad = tmp_path / "bad.qb3"
    bad.write_bytes(b"not a qb3 stream at all")
    out = subprocess.run([exe, "-i", str(bad)], capture_output=True, text=True)
    assert out.returncode != 0 and "error" in out.stdout
    assert subprocess.run([exe], capture_output=True, text=True).returncode != 0  # usage


@pytest.mark.skipif(not os.environ.get("QB3_TEST_CMAKE") or not shutil.which("cmake"), reason="set QB3_TEST_CMAKE=1: rebuilds all kernels")
def test_cmake_install_and_find_package(tmp_path):
    b, inst, c = str(tmp_path / "b"), str(tmp_path / "inst"), str(tmp_path / "c")
    subprocess.run(["cmake", "-S", ROOT, "-B", b, "-DCMAKE_BUILD_TYPE=Release"], check=True, capture_output=True)
    subprocess.run(["cmake", "--build", b, "-j4"], check=True, capture_output=True)
    subprocess.run(["cmake", "--install", b, "--prefix", inst], check=True, capture_output=True)
    assert os.path.exists(os.path.join(inst, "lib", "libQB3.so")) and os.path.exists(os.path.join(inst, "include", "QB3.h"))
    subprocess.run(["cmake", "-S", os.path.join(ROOT, "cmake", "consumer"), "-B", c, "-DCMAKE_PREFIX_PATH=" + inst], check=True, capture_output=True)
    subprocess.run(["cmake", "--build", c], check=True, capture_output=True)
    assert subprocess.run([os.path.join(c, "consumer")], capture_output=True, text=True).stdout.strip() == "QB3 package ok"
