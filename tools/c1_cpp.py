"""Config C1 timed through the drop-in C++ CLASSES (not the Python mirror): ORB_SLAM2::ORBextractor::operator() on two 640x480
frames (tests/cpp/c1_main, built against the test-only cv shim) + ORB_SLAM2::ORBmatcher::SearchForInitialization of the
reference's own ORBmatcher class with the drop-in bodies (oracle/_ref/libfbe_dropinmatch.so; the method call alone is timed
inside the harness).  Everything is checked against the oracle.  Prints one JSON line.
   python tools/c1_cpp.py [reps]"""
import ctypes as C
import json
import os
import struct
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from fishbirdeyevisualslam_b200 import synth
from fishbirdeyevisualslam_b200._lib import KP_DTYPE
from fishbirdeyevisualslam_b200.matcher import Frame
from oracle import oracle as O

HOST = os.path.join(ROOT, "fishbirdeyevisualslam_b200", "host")
EXE = os.path.join(ROOT, "tests", "cpp", "c1_main")


def build_exe():
    pkg = os.path.join(ROOT, "fishbirdeyevisualslam_b200")
    subprocess.run(["g++", "-std=c++11", "-O2", "-I", os.path.join(ROOT, "oracle", "cvshim"), "-I", HOST, "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "c1_main.cc"), os.path.join(HOST, "ORBextractor.cc"),
                    os.path.join(ROOT, "oracle", "cvshim", "shim.cpp"), "-L", pkg, "-lfbe_b200", f"-Wl,-rpath,{pkg}", "-o", EXE], check=True)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    build_exe()
    h, w = 480, 640
    a, b = synth.frame_pair_in_time(h, w, 1001)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        pa, pb, po = (os.path.join(td, n) for n in ("a.raw", "b.raw", "out.bin"))
        a.tofile(pa); b.tofile(pb)
        for tag, env in (("with_image_pyramid", {}), ("no_image_pyramid", {"FBE_IMAGE_PYRAMID": "0"})):
            r = subprocess.run([EXE, str(h), str(w), "1000", "8", pa, pb, po, str(reps)], check=True, capture_output=True, text=True,
                               env=dict(os.environ, **env))
            out["extract_x2_us_" + tag] = float(r.stdout.strip())
        raw = open(po, "rb").read()
    frames, off = [], 0
    for _ in range(2):
        n = struct.unpack_from("<i", raw, off)[0]
        k = np.frombuffer(raw, KP_DTYPE, n, off + 4).copy()
        d = np.frombuffer(raw, np.uint8, n * 32, off + 4 + n * 28).reshape(n, 32).copy()
        frames.append((k, d))
        off += 4 + n * 60
    oe = O.OracleExtractor(1000, 1.2, 8, 15, 5)
    (ka, da), (kb, db) = frames
    oka, oda = oe(a)
    okb, odb = oe(b)
    parity = ka.tobytes() == oka.tobytes() and np.array_equal(da, oda) and kb.tobytes() == okb.tobytes() and np.array_equal(db, odb)
    F1, F2 = Frame.front(ka, da, w, h), Frame.front(kb, db, w, h)
    D = O.dropinmatch()
    D.refm_last_method_us.restype = C.c_double
    rm = O.RefMatch(D)
    ts = []
    for r in range(reps + 3):
        pm = np.ascontiguousarray(np.stack([ka["x"], ka["y"]], 1), np.float32)
        n, m12 = rm.search_for_initialization(F1, F2, pm, 100, 0.9, True)
        if r >= 3:
            ts.append(D.refm_last_method_us())
    pmo = np.ascontiguousarray(np.stack([ka["x"], ka["y"]], 1), np.float32)
    no, m12o = O.search_for_initialization(F1, F2, pmo, 100, 0.9, True)
    parity = parity and n == no and np.array_equal(m12, m12o) and np.array_equal(pm, pmo)
    out["search_for_initialization_us"] = float(np.median(ts))
    out["c1_ms_per_pair_cpp_classes"] = (out["extract_x2_us_with_image_pyramid"] + out["search_for_initialization_us"]) / 1e3
    out["c1_ms_per_pair_cpp_classes_no_image_pyramid"] = (out["extract_x2_us_no_image_pyramid"] + out["search_for_initialization_us"]) / 1e3
    out.update({"config": "C1 640x480 pair @1000: ORBextractor::operator() x2 + ORBmatcher::SearchForInitialization(window 100), drop-in C++ classes",
                "parity": bool(parity), "keypoints": [len(ka), len(kb)], "matches": int(n), "reps": reps})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
