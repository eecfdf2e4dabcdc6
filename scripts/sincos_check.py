"""Development aid: exhaustive (all 2^32 floats) checks on the GPU of two identities the kernels could use:
sincosf(a) == (sinf(a), cosf(a)) bit for bit, and powf(a, 2.0f) == a * a bit for bit."""
import sys, ctypes as C
sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import _lib
lib = _lib.load()
m = (C.c_ulonglong * 3)()
f = (C.c_uint * 3)()
rc = lib.gmt_debug_sincos_check(0, m, f)
print("rc", rc)
print("sincosf vs sinf/cosf: mismatching floats:", m[0], "first bad bits: 0x%08x" % f[0])
print("powf(a, 2) vs a * a : mismatching floats:", m[1], "first bad bits: 0x%08x" % f[1])
print("sincos_small vs sincosf (|a| < 105615): mismatching floats:", m[2], "first bad bits: 0x%08x" % f[2])
