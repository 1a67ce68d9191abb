"""The packed-record variants of the fused closures (pair_m / pair_rec, csrc/wcsph_ops.cuh) restate
pair() with the neighbour's fields taken from its records instead of the SoA arrays.  The two
bodies must stay the same arithmetic, token for token, or the record path would no longer be
bit-identical (wcsph_perturbed_witch.jl:226-228, :261-286).  This test compares the source text
after the substitution; it needs no GPU."""
import re
from pathlib import Path

import pytest

SRC = (Path(__file__).resolve().parent.parent / "sph_mountain_waves_b200" / "csrc" / "wcsph_ops.cuh").read_text()

# field of q in the SoA closure -> the same value in the record closure
SUBST = [
    (r"jl_max\(QF\(S_RHO\), c\.rho_floor\)", "C.b"),
    (r"QF\(S_V0\)", "B.a"), (r"QF\(S_V1\)", "B.b"), (r"QF\(S_V2\)", "B.c"), (r"QF\(S_H\)", "B.d"),
    (r"QF\(S_PR2\)", "C.a"), (r"QF\(S_CS\)", "C.c"), (r"QF\(S_M\)", "qm"),
]


def struct_body(name: str) -> str:
    m = re.search(r"struct %s : PairOpBase \{" % re.escape(name), SRC)
    assert m, name
    depth, i = 1, m.end()
    while depth:
        depth += {"{": 1, "}": -1}.get(SRC[i], 0)
        i += 1
    return SRC[m.end():i - 1]


def method_body(struct_src: str, method: str) -> str:
    m = re.search(r"__device__ void %s\(" % method, struct_src)
    assert m, method
    i = struct_src.index("{", m.end())
    depth, j = 1, i + 1
    while depth:
        depth += {"{": 1, "}": -1}.get(struct_src[j], 0)
        j += 1
    return struct_src[i + 1:j - 1]


def normalise(body: str) -> str:
    body = re.sub(r"//[^\n]*", "", body)
    body = re.sub(r"double qm = qm;", "", body)          # the record closure gets qm as an argument
    return re.sub(r"\s+", " ", body).strip()


@pytest.mark.parametrize("struct,record_method", [
    ("B_wcsph_density_fused", "pair_m"), ("B_wcsph_density_fast", "pair_m"),
    ("B_wcsph_momentum_fused", "pair_rec"), ("B_wcsph_momentum_fast", "pair_rec")])
def test_record_closure_is_the_same_arithmetic(struct, record_method):
    s = struct_body(struct)
    soa = method_body(s, "pair")
    for pat, rep in SUBST:
        soa = re.sub(pat, rep, soa)
    assert "QF(" not in soa, "a neighbour field without a record slot: " + soa
    assert normalise(soa) == normalise(method_body(s, record_method))
