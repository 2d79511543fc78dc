"""yuvdiff.py a.yuv b.yuv width height -- first differing pictures / planes / 4x4 units of two 10-bit 4:2:0 files (debugging aid)"""
import sys
import numpy as np

a, b, w, h = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
A, B = np.fromfile(a, np.uint16), np.fromfile(b, np.uint16)
fs = w * h * 3 // 2
print("frames", len(A) / fs, len(B) / fs)
for f in range(min(len(A), len(B)) // fs):
    o = f * fs
    for name, off, pw, ph in (("Y", 0, w, h), ("Cb", w * h, w // 2, h // 2), ("Cr", w * h * 5 // 4, w // 2, h // 2)):
        x = A[o + off:o + off + pw * ph].reshape(ph, pw).astype(int)
        y = B[o + off:o + off + pw * ph].reshape(ph, pw).astype(int)
        d = x != y
        if d.any():
            ys, xs = np.nonzero(d)
            u = sorted(set(zip((ys // 4).tolist(), (xs // 4).tolist())))
            print("frame %d %s: %d samples differ, %d 4x4 units; first (x,y) = (%d,%d) a=%d b=%d; bbox x %d..%d y %d..%d; max |d| %d"
                  % (f, name, d.sum(), len(u), xs[0], ys[0], x[ys[0], xs[0]], y[ys[0], xs[0]], xs.min(), xs.max(), ys.min(), ys.max(), np.abs(x - y).max()))
            if len(sys.argv) > 5:
                print("   units (x,y):", [(c * 4, r * 4) for r, c in u[:40]])
    if len(sys.argv) <= 5 and (A[o:o + fs] != B[o:o + fs]).any():
        break
