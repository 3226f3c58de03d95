"""Julia is not installed in the build image, so julia/AGPBlackwell.jl cannot be executed here.  This static check keeps
its `ccall`s honest against the tested ctypes mirror: every symbol exists in include/agp.h / _cabi.SIGNATURES and is
called with the same NUMBER of arguments, with pointer-vs-integer kinds in the same positions."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def _ccalls(src):
    for m in re.finditer(r"ccall\(\(:(\w+), libagp\),\s*(\w+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        yield m.group(1), m.group(2), _split_top(src[m.end():i - 1])


def test_every_ccall_matches_the_ctypes_mirror():
    import agp_b200
    sigs = agp_b200._cabi.SIGNATURES
    src = open(os.path.join(ROOT, "julia", "AGPBlackwell.jl")).read()
    seen = set()
    for name, ret, types in _ccalls(src):
        assert name in sigs, name
        res, args = sigs[name]
        assert len(types) == len(args), (name, types, args)
        for jt, ct in zip(types, args):
            is_ptr_j = jt.startswith(("Ptr", "Ref", "Cstring"))
            is_ptr_c = ct in (C.c_void_p, C.c_char_p) or hasattr(ct, "contents") or ct.__name__.startswith("LP_")
            assert is_ptr_j == is_ptr_c, (name, jt, ct)
            if not is_ptr_j:
                assert {"Int32": C.c_int32, "Int64": C.c_int64, "Float64": C.c_double}[jt] is ct, (name, jt, ct)
        seen.add(name)
    # the hot-path entry points SURVEY s8(b) lists are all bound
    for need in ("agp_init", "agp_fit", "agp_post_mean_var", "agp_post_solve_lower", "agp_post_factor_export", "agp_post_extend",
                 "agp_post_logpdf", "agp_post_rand", "agp_rand", "agp_vfe_elbo", "agp_vfe_fit", "agp_vfe_mean_var",
                 "agp_post_free", "agp_vfe_post_free", "agp_post_logpdf_grad", "agp_last_error"):
        assert need in seen, need
