"""Java `Double.toString` layout, so emitResult prints what the reference's string interpolation prints
(VariantsPca.scala:239, :243)."""
from __future__ import annotations

import math


def jdouble(x: float) -> str:
    """Plain decimal for 1e-3 <= |x| < 1e7, otherwise d.dddE[-]n; digits are the shortest that round-trip
    (what JDK >= 19 prints; older JDKs occasionally print one digit more)."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0.0:
        return "-0.0" if math.copysign(1.0, x) < 0 else "0.0"
    sign = "-" if x < 0 else ""
    r = repr(abs(x))
    if "e" in r:
        mant, e = r.split("e")
        exp = int(e)
    else:
        mant = r
        exp = int(f"{abs(x):.17e}".split("e")[1])
    digits = mant.replace(".", "").lstrip("0").rstrip("0") or "0"
    if 1e-3 <= abs(x) < 1e7:
        if exp >= 0:
            ip, fp = digits[: exp + 1].ljust(exp + 1, "0"), digits[exp + 1:]
        else:
            ip, fp = "0", "0" * (-exp - 1) + digits
        return f"{sign}{ip}.{fp or '0'}"
    return f"{sign}{digits[0]}.{digits[1:] or '0'}E{exp}"
