#!/usr/bin/env python
"""Known-answer vectors for the WAH codec (tests/golden/wah_vectors.json), produced by an implementation that
shares NOTHING with oracle/wah_oracle.c or the product writer: pure Python over lists of bits, written from the
published description of the format (Wu / Otoo / Shoshani, TODS 31(1) 2006, §3: 31-bit groups, literal words with
MSB 0 and the first bit most significant, fill words MSB 1 / bit 30 = fill bit / low 30 bits = groups, the trailing
n mod 31 bits kept apart as the active word) and from the append rule of FastBit's writer as published in its
bitvector.h (ibis::bitvector::append_active): a single all-zero or all-one group is appended as a LITERAL word; the
second one in a row turns that literal into a fill of 2; later ones increment the fill.  Neither FastBit nor the
CUBIT library is in /root/reference, so this — with the paper's own 128-bit example, which is vector 0 —
is the pin available.

Run: python tests/golden/make_wah_golden.py   (deterministic; rewrites wah_vectors.json)
"""
import json
import os
import random

ALLONES = 0x7FFFFFFF
HEADER0, HEADER1 = 0x80000000, 0xC0000000


def wah_encode(bits):
    words = []
    i = 0
    while len(bits) - i >= 31:
        val = 0
        for b in bits[i:i + 31]:
            val = (val << 1) | b
        i += 31
        if val == 0:
            if words and words[-1] == 0:
                words[-1] = HEADER0 + 2
            elif words and HEADER0 <= words[-1] < HEADER1:
                words[-1] += 1
            else:
                words.append(0)
        elif val == ALLONES:
            if words and words[-1] == ALLONES:
                words[-1] = HEADER1 | 2
            elif words and words[-1] >= HEADER1:
                words[-1] += 1
            else:
                words.append(ALLONES)
        else:
            words.append(val)
    rest = bits[i:]
    av = 0
    for b in rest:
        av = (av << 1) | b
    return words, av, len(rest)


def expand(runs):
    bits = []
    for bit, n in runs:
        bits += [bit] * n
    return bits


def main():
    rnd = random.Random(20060331)
    cases = [
        ("paper_fig1", [(1, 1), (0, 20), (1, 3), (0, 79), (1, 25)]),
        ("one_bit", [(1, 1)]),
        ("thirty_zeros", [(0, 30)]),
        ("one_zero_group", [(0, 31)]),                      # a lone fill group stays a literal 00000000
        ("one_one_group", [(1, 31)]),                       # ... 7FFFFFFF
        ("two_zero_groups", [(0, 62)]),                     # literal → fill of 2
        ("two_one_groups_and_a_bit", [(1, 63)]),
        ("three_one_groups", [(1, 93)]),
        ("zero_one_zero_groups", [(0, 31), (1, 31), (0, 31)]),   # three literals, no fill
        ("fill_then_lone_group", [(0, 31 * 5), (1, 31), (0, 31), (1, 31 * 2), (0, 7)]),
        ("literal_between_fills", [(0, 31 * 4), (1, 1), (0, 30), (0, 31 * 3)]),
        ("fill_boundary_off_by_one", [(0, 31 * 3 - 1), (1, 1), (0, 31 * 3 + 1), (1, 30)]),
        ("long_zero_fill", [(0, 31 * 70000), (1, 5)]),
        ("long_one_fill", [(1, 31 * 70001)]),
        ("alternating_bits", [(i & 1, 1) for i in range(200)]),
        ("exact_multiple_no_active", [(1, 10), (0, 21), (0, 31), (0, 31), (1, 31)]),
    ]
    for c in range(24):   # seeded random run lists: short and long runs mixed, lengths not aligned to groups
        runs, bit = [], rnd.randint(0, 1)
        for _ in range(rnd.randint(1, 40)):
            kind = rnd.random()
            n = rnd.randint(1, 6) if kind < 0.4 else rnd.randint(20, 140) if kind < 0.8 else rnd.randint(200, 5000)
            runs.append((bit, n))
            bit ^= 1
        cases.append(("random_%02d" % c, runs))
    out = []
    for name, runs in cases:
        words, av, an = wah_encode(expand(runs))
        out.append({"name": name, "runs": [list(r) for r in runs], "n_bits": sum(n for _, n in runs),
                    "wah": ["%08X" % w for w in words], "active_val": "%08X" % av, "active_nbits": an})
    assert out[0]["wah"] == ["40000380", "80000002", "001FFFFF"] and out[0]["active_val"] == "0000000F"
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wah_vectors.json")
    with open(path, "w") as f:
        json.dump({"source": "tests/golden/make_wah_golden.py (independent pure-Python encoder)", "vectors": out}, f, indent=0)
    print("wrote %d vectors to %s" % (len(out), path))


if __name__ == "__main__":
    main()
