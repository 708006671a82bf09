#!/usr/bin/env python3
"""Host check of research/dfma.cu's mul52 dump: out == a*b*2^-260 (mod q), out < 2q, limbs < 2^52."""
import struct, sys
Q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
data = open(sys.argv[1], "rb").read()
n = 4096
vals = struct.unpack("<%dQ" % (n * 15), data)
rinv = pow(1 << 260, -1, Q)
bad = 0
for t in range(n):
    a = sum(vals[t * 10 + i] << (52 * i) for i in range(5))
    b = sum(vals[t * 10 + 5 + i] << (52 * i) for i in range(5))
    o = vals[n * 10 + t * 5: n * 10 + t * 5 + 5]
    ov = sum(o[i] << (52 * i) for i in range(5))
    if any(x >> 52 for x in o[:4]) or ov >= 2 * Q or ov % Q != a * b * rinv % Q:
        bad += 1
        if bad < 4:
            print("mismatch at", t, hex(a), hex(b), hex(ov))
print("mul52 check:", "OK" if not bad else f"{bad} BAD", "of", n)
