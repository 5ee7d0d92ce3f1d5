"""TEST INFRASTRUCTURE. Locate OpenCV's LineAA / ellipse lookup tables inside the installed cv2 wheel's
.rodata (icvSlopeCorrTable, FilterTable, SinTable of modules/imgproc/src/drawing.cpp) and print them,
so oracle/cvdraw_tables.h can be checked against the binary the parity tests run against.

    python oracle/tools/extract_cv_tables.py            # prints the three tables
"""
import os
import numpy as np


def extract():
    import cv2
    so = os.path.join(os.path.dirname(cv2.__file__), "cv2.abi3.so")
    data = open(so, "rb").read()
    i = data.find(bytes([181, 181, 181, 182, 182, 183, 184, 185]))
    j = data.find(np.array([168, 177, 185, 194], dtype=np.int32).tobytes())
    k = data.find(np.array([0.0, 0.0174524, 0.0348995], dtype=np.float32).tobytes())
    if min(i, j, k) < 0:
        raise RuntimeError("tables not found in " + so)
    slope = list(data[i:i + 32])
    filt = np.frombuffer(data[j:j + 256], dtype=np.int32).tolist()
    sin = np.frombuffer(data[k:k + 4 * 451], dtype=np.float32).copy()
    return slope, filt, sin


if __name__ == "__main__":
    s, f, t = extract()
    print("slope_corr", s)
    print("filter", f)
    print("sin[0:8]", t[:8].tolist(), "... sin == float32(round(sin(deg),7)):",
          all(np.float32(round(float(np.sin(np.deg2rad(d))), 7)) == t[d] for d in range(451)))
