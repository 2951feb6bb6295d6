"""Floating-point clamp constants of the probability scales (fp64 only).

Same values as the reference's probayes/constants.py:9-32; the device path is
fp64 throughout, so only the 64-bit set exists."""
import math

NEARLY_POSITIVE_ZERO = 2.2250738585072014e-308
NEARLY_POSITIVE_INF = 1.7976931348623158e+308
NEARLY_NEGATIVE_INF = -NEARLY_POSITIVE_INF
LOG_NEARLY_POSITIVE_INF = math.log(NEARLY_POSITIVE_INF)
COMPLEX_ZERO = complex(0., 0.)
