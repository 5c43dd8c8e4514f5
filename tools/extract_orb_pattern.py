"""Reads ORB's 256 x 4 sampling pattern (`bit_pattern_31_`, modules/features2d/src/orb.cpp: the learned BRIEF test
positions for patch size 31) out of the installed cv2 binary and writes epivo_b200/orb_pattern.py and
epivo_b200/csrc/orb_pattern.inc.  The pattern is data of the OpenCV dependency, not reference source; the table's first
and last published rows (8,-3, 9,5 ... -1,-6, 0,-11) locate and check it.

    python tools/extract_orb_pattern.py
"""
import glob
import os

import numpy as np


def main():
    import cv2
    so = glob.glob(os.path.join(os.path.dirname(cv2.__file__), "*.so"))[0]
    data = open(so, "rb").read()
    key = np.array([8, -3, 9, 5, 4, 2, 7, -12, -11, 9, -8, 2], dtype="<i4").tobytes()
    at = data.find(key)
    assert at >= 0 and data.find(key, at + 1) < 0, "pattern not found exactly once"
    pat = np.frombuffer(data[at:at + 4096], dtype="<i4").reshape(256, 4)
    assert pat[-1].tolist() == [-1, -6, 0, -11] and np.abs(pat).max() <= 13
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rows = ["    " + ", ".join("%d" % v for v in r) + "," for r in pat]
    with open(os.path.join(root, "epivo_b200", "orb_pattern.py"), "w") as f:
        f.write('"""ORB sampling pattern for patch size 31 (OpenCV orb.cpp bit_pattern_31_): 256 tests x (x0, y0, x1, y1).\n'
                'Written by tools/extract_orb_pattern.py from cv2 %s."""\n' % cv2.__version__)
        f.write("import numpy as np\n\nBIT_PATTERN_31 = np.array([\n" + "\n".join(rows) + "\n], dtype=np.int8).reshape(256, 4)\n")
    with open(os.path.join(root, "epivo_b200", "csrc", "orb_pattern.inc"), "w") as f:
        f.write("// ORB sampling pattern for patch size 31 (OpenCV orb.cpp bit_pattern_31_): 256 tests x (x0, y0, x1, y1).\n"
                "// Written by tools/extract_orb_pattern.py from cv2 %s.\n" % cv2.__version__)
        f.write("\n".join(rows) + "\n")
    print("wrote pattern,", pat.shape)


if __name__ == "__main__":
    main()
