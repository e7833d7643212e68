"""Distance transforms: the closed set of four that can cross the C ABI.

Mirrors the constructors of the reference (src/transforms/distancetransforms.jl:16-25).
There a transform is a *string* eval'd into a Julia closure, and the string is what is stored
and serialised (src/transforms/lambdas.jl:9-12,46-51).  A closure cannot cross a C ABI, so this
module keeps the same four constructors, the same source strings (``exstr``), and adds a parser
that maps an ``exstr`` back to ``(kind, params)``.  Anything else is rejected loudly, which is
the documented behaviour of the boundary.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass

# kind codes shared with include/aceb200.h
TRANS_ID, TRANS_POLY, TRANS_MORSE, TRANS_AGNESI = 0, 1, 2, 3


def _jl(x) -> str:
    """Decimal printout Julia's string interpolation would produce for a number."""
    if isinstance(x, bool):
        raise TypeError("bool is not a transform parameter")
    if isinstance(x, int):
        return str(x)
    r = repr(float(x))
    return r


@dataclass(frozen=True)
class Lambda:
    """A distance transform with its reference source string (lambdas.jl:9-12)."""

    kind: int
    params: tuple  # see aceb200.h: poly (p, r0); morse (lambda, r0); agnesi (r0, p, a)
    exstr: str

    # value t(r): distancetransforms.jl:16-25
    def __call__(self, r: float) -> float:
        k, q = self.kind, self.params
        if k == TRANS_ID:
            return r
        if k == TRANS_POLY:
            p, r0 = q
            return ((1.0 + r0) / (1.0 + r)) ** p
        if k == TRANS_MORSE:
            lam, r0 = q
            return math.exp(-lam * (r / r0 - 1.0))
        if k == TRANS_AGNESI:
            r0, p, a = q
            return 1.0 / (1.0 + a * (r / r0) ** p)
        raise ValueError("unknown transform kind")

    # derivative t'(r): the reference uses ForwardDiff (lambdas.jl:28-34); these are the closed forms
    def deriv(self, r: float) -> float:
        k, q = self.kind, self.params
        if k == TRANS_ID:
            return 1.0
        if k == TRANS_POLY:
            p, r0 = q
            return -p * ((1.0 + r0) / (1.0 + r)) ** p / (1.0 + r)
        if k == TRANS_MORSE:
            lam, r0 = q
            return -(lam / r0) * math.exp(-lam * (r / r0 - 1.0))
        if k == TRANS_AGNESI:
            r0, p, a = q
            x = r / r0
            d = 1.0 + a * x ** p
            return -a * p * x ** (p - 1) / r0 / (d * d)
        raise ValueError("unknown transform kind")

    def inv(self, t: float) -> float:
        """Inverse transform; the reference uses a root finder (distancetransforms.jl:32-38)."""
        k, q = self.kind, self.params
        if k == TRANS_ID:
            return t
        if k == TRANS_POLY:
            p, r0 = q
            return (1.0 + r0) / t ** (1.0 / p) - 1.0
        if k == TRANS_MORSE:
            lam, r0 = q
            return r0 * (1.0 - math.log(t) / lam)
        if k == TRANS_AGNESI:
            r0, p, a = q
            return r0 * ((1.0 / t - 1.0) / a) ** (1.0 / p)
        raise ValueError("unknown transform kind")

    def c_params(self):
        q = list(self.params) + [0.0] * (4 - len(self.params))
        return [float(v) for v in q]


def polytransform(p, r0) -> Lambda:
    return Lambda(TRANS_POLY, (float(p), float(r0)), f"r -> ((1+{_jl(r0)})/(1+r))^{_jl(p)}")


def idtransform() -> Lambda:
    return Lambda(TRANS_ID, (), "r -> r")


def morsetransform(lam, r0) -> Lambda:
    return Lambda(TRANS_MORSE, (float(lam), float(r0)), f"r -> exp(- {_jl(lam)} * (r / {_jl(r0)} - 1))")


def agnesitransform(r0, p=2, a=None) -> Lambda:
    if a is None:
        a = (p - 1) / (p + 1)
    return Lambda(TRANS_AGNESI, (float(r0), float(p), float(a)),
                  f"r -> 1 / (1 + {_jl(a)} * (r / {_jl(r0)})^{_jl(p)})")


_NUM = r"([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+))"
_PATTERNS = [
    (TRANS_ID, re.compile(r"^r->r$"), lambda m: ()),
    (TRANS_POLY, re.compile(rf"^r->\(\(1\+{_NUM}\)/\(1\+r\)\)\^{_NUM}$"),
     lambda m: (float(m.group(2)), float(m.group(1)))),
    (TRANS_MORSE, re.compile(rf"^r->exp\(-{_NUM}\*\(r/{_NUM}-1\)\)$"),
     lambda m: (float(m.group(1)), float(m.group(2)))),
    (TRANS_AGNESI, re.compile(rf"^r->1/\(1\+{_NUM}\*\(r/{_NUM}\)\^{_NUM}\)$"),
     lambda m: (float(m.group(2)), float(m.group(3)), float(m.group(1)))),
]


def parse_exstr(exstr: str) -> Lambda:
    """Map a serialised ``Lambda.exstr`` (lambdas.jl:46-51) to one of the four supported kinds.

    Raises ``ValueError`` for any other lambda: an arbitrary Julia closure cannot run on the GPU.
    """
    s = exstr.replace(" ", "")
    for kind, pat, conv in _PATTERNS:
        m = pat.match(s)
        if m:
            return Lambda(kind, conv(m), exstr)
    raise ValueError(f"unsupported distance transform for the B200 path: {exstr!r} "
                     "(supported: idtransform, polytransform, morsetransform, agnesitransform)")
