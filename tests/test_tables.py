"""The committed case tables are the reference's topology specification:
check the values against the sha256 prefixes recorded in the survey, without needing
/root/reference at test time."""
import hashlib
import os
import re
import struct

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EXPECT = {
    "T_MC_CORNERS": (24, "c4d311fec9c59731"),
    "T_DMC_CORNERS": (24, "1b7d213f5218b834"),
    "T_EDGE_LOC": (48, "96abbef0a3fa9825"),
    "T_MC_FIRST": (257, "928c98cc7322ed23"),
    "T_MC_IDS": (2460, "7d0cab9e345c9280"),
    "T_PROBLEMATIC": (256, "700dec29e174660d"),
    "T_PATCH_FIRST": (257, "535b3360a6fe996d"),
    "T_EDGE_FIRST": (359, "a38a1b59c13ddf3e"),
    "T_EDGE_INDEX": (1536, "17a15fccec6fb3a0"),
    "T_DMC_EDGE_OFFSET": (3072, "b5ed1a01713244ad"),
    "T_DMC_QUAD": (96, "2f295da9d5a54215"),
}


def _parse(path):
    src = open(path).read()
    out = {}
    for m in re.finditer(r"(\w+)\[(\d+)\] = \{(.*?)\};", src, re.S):
        out[m.group(1)] = [int(x, 0) for x in re.findall(r"-?(?:0x[0-9a-fA-F]+|\d+)", m.group(3))]
    return out


def test_oracle_tables_match_survey_checksums():
    t = _parse(os.path.join(ROOT, "oracle", "diso_tables.h"))
    for name, (n, want) in EXPECT.items():
        vals = t[name]
        assert len(vals) == n, name
        got = hashlib.sha256(struct.pack("<%di" % n, *vals)).hexdigest()[:16]
        assert got == want, name


def test_packed_kernel_tables_consistent_with_oracle_tables():
    t = _parse(os.path.join(ROOT, "oracle", "diso_tables.h"))
    k = _parse(os.path.join(ROOT, "diso_b200", "csrc", "case_tables.inc"))
    mf, mi = t["T_MC_FIRST"], t["T_MC_IDS"]
    for code in range(256):
        ids = mi[mf[code]:mf[code + 1]]
        w = k["T_MC_CASE"][code]
        assert (w >> 60) == len(ids) // 3
        assert [(w >> (4 * i)) & 15 for i in range(len(ids))] == ids
    pf, ef, ei = t["T_PATCH_FIRST"], t["T_EDGE_FIRST"], t["T_EDGE_INDEX"]
    off, prob = t["T_DMC_EDGE_OFFSET"], t["T_PROBLEMATIC"]
    for code in range(256):
        w, c = k["T_DMC_CASE"][code], k["T_DMC_PATCHLEN"][code]
        npatch = pf[code + 1] - pf[code]
        assert (w >> 24) & 7 == npatch
        for e in range(12):
            o = off[code * 12 + e]
            assert ((c >> (16 + e)) & 1) == (o >= 0)
            if o >= 0:
                assert (w >> (2 * e)) & 3 == o
        for p in range(npatch):
            assert (c >> (4 * p)) & 15 == ef[pf[code] + p + 1] - ef[pf[code] + p]
        mem = k["T_DMC_MEMBERS"][code]
        for q in range(4):
            want_m = sum(1 << e for e in range(12) if off[code * 12 + e] == q)
            assert (mem >> (12 * q)) & 0xfff == want_m
        assert mem >> 48 == c & 0xffff      # the patch lengths ride in the top 16 bits
        assert (w >> 31) == (prob[code] != 255)
        if prob[code] != 255:
            assert (w >> 28) & 7 == prob[code]


def test_table_structure_assumptions():
    """Facts the kernels rely on (also asserted by tools/extract_tables.py at generation)."""
    t = _parse(os.path.join(ROOT, "oracle", "diso_tables.h"))
    pf, ef, ei = t["T_PATCH_FIRST"], t["T_EDGE_FIRST"], t["T_EDGE_INDEX"]
    for p in range(358):
        edges = ei[ef[p]:ef[p + 1]]
        assert edges == sorted(edges)  # dual-vertex summation order == ascending edge id
    loc = t["T_EDGE_LOC"]
    dx = sum(loc[4 * e] << e for e in range(12))
    dy = sum(loc[4 * e + 1] << e for e in range(12))
    dz = sum(loc[4 * e + 2] << e for e in range(12))
    ax = sum(loc[4 * e + 3] << (2 * e) for e in range(12))
    src = open(os.path.join(ROOT, "diso_b200", "csrc", "tables.cuh")).read()
    for name, val in (("EDGE_DX", dx), ("EDGE_DY", dy), ("EDGE_DZ", dz), ("EDGE_AX", ax)):
        m = re.search(name + r" = (0x[0-9a-f]+)u", src)
        assert int(m.group(1), 16) == val, name


def test_decoded_triangle_table_consistent_with_oracle_tables():
    """T_MC_TRI5 (the triangle kernel's pre-decoded corner codes) against the triangle list and the edge geometry table
    (mcEdgeLocations, cumc.cu:109-122: per local edge {dx, dy, dz, axis} of the owning point)."""
    t = _parse(os.path.join(ROOT, "oracle", "diso_tables.h"))
    k = _parse(os.path.join(ROOT, "diso_b200", "csrc", "case_tables.inc"))
    mf, mi, loc, tri5 = t["T_MC_FIRST"], t["T_MC_IDS"], t["T_EDGE_LOC"], k["T_MC_TRI5"]
    assert len(tri5) == 1024
    for code in range(256):
        ids = mi[mf[code]:mf[code + 1]]
        assert tri5[4 * code + 3] == len(ids) // 3
        for i, e in enumerate(ids):
            dx, dy, dz, ax = loc[4 * e:4 * e + 4]
            q, c = divmod(i, 3)
            f = (tri5[4 * code + (q >> 1)] >> (15 * (q & 1) + 5 * c)) & 31
            assert f == (2 * dx + dy) | (dz << 2) | ((ax >= 1) << 3) | ((ax == 2) << 4), (code, i, e)
        used = 5 * len(ids)     # nothing stored beyond the case's triangles
        for w in range(3):
            bits = max(0, min(30, used - 30 * w))
            assert tri5[4 * code + w] >> bits == 0


def test_dmc_edge5_table_consistent_with_packed_tables():
    """T_DMC_EDGE5 = per edge {length of its patch : 3, index of the patch in the cell : 2}, against T_DMC_CASE / T_DMC_PATCHLEN
    (themselves checked against the oracle tables above)."""
    k = _parse(os.path.join(ROOT, "diso_b200", "csrc", "case_tables.inc"))
    for code in range(256):
        w, c, e5 = k["T_DMC_CASE"][code], k["T_DMC_PATCHLEN"][code], k["T_DMC_EDGE5"][code]
        for e in range(12):
            f = (e5 >> (5 * e)) & 31
            if (c >> (16 + e)) & 1:
                o = (w >> (2 * e)) & 3
                assert f == ((c >> (4 * o)) & 15) | (o << 3)
            else:
                assert f == 0
        assert e5 >> 60 == 0
