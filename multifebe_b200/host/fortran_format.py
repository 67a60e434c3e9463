"""Fortran edit descriptors the reference's export files are written with (src/read_export.f90:61-221).

The result files (*.nso ...) are the reference's interface towards post-processing scripts, so their numbers must come out
character for character as gfortran prints them: `iW`, `eW.DeE` (0.ddd x 10^n) and `enW.DeE` (engineering notation: exponent
a multiple of three, 1 <= |mantissa| < 1000).  Rounding is done on the exact decimal expansion of the binary value (round half
to even, gfortran's default RN mode on an exactly representable tie), with `decimal`.
"""
import decimal as _d
import re

# named formats of the [export] section (src/read_export.f90:164-170) and the default (:66)
REAL_FORMATS = {"sci_double": "e25.16e3", "sci_simple": "e16.8e2", "sci_less": "e11.3e2", "eng_double": "en27.16e3",
                "eng_simple": "en18.8e2", "eng_less": "en13.3e2", "auto": "en27.16e3"}
DEFAULT_REAL_FORMAT = "en18.8e2"

_CTX = _d.Context(prec=400, rounding=_d.ROUND_HALF_EVEN)


def parse_real_format(fmt):
    """'en18.8e2' -> ('en', 18, 8, 2)."""
    fmt = REAL_FORMATS.get(fmt.strip().lower(), fmt.strip().lower())
    m = re.fullmatch(r"(en|es|e)(\d+)\.(\d+)e(\d+)", fmt)
    if not m:
        raise ValueError("unsupported real format %r (e, es or en edit descriptors with an explicit exponent width)" % fmt)
    return m.group(1), int(m.group(2)), int(m.group(3)), int(m.group(4))


def fmt_int(n, w):
    s = "%d" % n
    return "*" * w if len(s) > w else s.rjust(w)


def int_width(*maxima):
    """`i` width of the result files: digits of the largest id or index + 1 (src/read_export.f90:61)."""
    return len("%d" % max(2, *maxima)) + 1


def _digits(x, exp10, d):
    """|x| / 10^exp10 rounded to d decimals, as an integer count of 10^-d units."""
    q = _CTX.divide(_d.Decimal(abs(x)), _d.Decimal(10) ** exp10)
    return int(_CTX.quantize(q * (_d.Decimal(10) ** d), _d.Decimal(1)))


def fmt_real(x, kind, w, d, e):
    """One value under `kind` in ('e', 'es', 'en') with width w, d decimals and e exponent digits."""
    x = float(x)
    if x != x or x in (float("inf"), float("-inf")):
        return ("NaN" if x != x else ("Infinity" if x > 0 else "-Infinity")).rjust(w)
    neg = x < 0.0 or (x == 0.0 and str(x).startswith("-"))
    if x == 0.0:
        exp10, units = 0, 0
    else:
        e10 = _d.Decimal(abs(x)).adjusted()           # floor(log10|x|)
        if kind == "e":
            exp10 = e10 + 1                           # mantissa in [0.1, 1)
        elif kind == "es":
            exp10 = e10                               # [1, 10)
        else:
            exp10 = 3 * (e10 // 3)                    # [1, 1000), exponent a multiple of 3
        units = _digits(x, exp10, d)
        top = {"e": 1, "es": 10, "en": 1000}[kind] * 10 ** d
        if units >= top:                              # rounding carried into the next decade
            exp10 += {"e": 1, "es": 1, "en": 3}[kind]
            units = _digits(x, exp10, d)
    ip, fp = divmod(units, 10 ** d)
    mant = "%d.%s" % (ip, ("%0*d" % (d, fp)) if d > 0 else "")
    ex = "E%s%0*d" % ("-" if exp10 < 0 else "+", e, abs(exp10))
    s = ("-" if neg else "") + mant + ex
    if len(s) > w and kind == "e" and mant.startswith("0."):
        s = ("-" if neg else "") + mant[1:] + ex      # the optional leading zero is dropped when the field is too narrow
    return "*" * w if len(s) > w else s.rjust(w)


class RealFormat:
    def __init__(self, fmt=DEFAULT_REAL_FORMAT):
        self.kind, self.w, self.d, self.e = parse_real_format(fmt)
        self.name = "%s%d.%de%d" % (self.kind, self.w, self.d, self.e)

    def __call__(self, x):
        return fmt_real(x, self.kind, self.w, self.d, self.e)
