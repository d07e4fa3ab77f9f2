import math


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def snr_db(a, b):
    return -20.0 * math.log10(max(rel_l2(a, b), 1e-30))
