"""Known-answer problems of the reference's levmar demo (levmar/lmdemo.c), restated as Python
callbacks.  Python floats are IEEE doubles and math.* calls the same libm, so with the expressions
written in the demo's order these callbacks return bit-identical values to the C ones.

Each entry: name, driver, m, n, p0, x, bounds, itmax, callbacks, and the solution/info the
reference prints (reproduced here with oracle/_ref).
"""
import math

import numpy as np

from oracle_lib import FUNC_T

DBL_MAX = 1.7976931348623157e308
# lmdemo.c:816-817
OPTS = (1e-3, 1e-15, 1e-15, 1e-20, 1e-6)


def _wrap(fn):
    def cb(p, out, m, n, _):
        fn(p, out, m, n)
    return FUNC_T(cb)


# lmdemo.c:47-67
def ros(p, x, m, n):
    for i in range(n):
        x[i] = ((1.0 - p[0]) * (1.0 - p[0]) + 105.0 * (p[1] - p[0] * p[0]) * (p[1] - p[0] * p[0]))


def jacros(p, jac, m, n):
    j = 0
    for i in range(n):
        jac[j] = (-2 + 2 * p[0] - 4 * 105.0 * (p[1] - p[0] * p[0]) * p[0]); j += 1
        jac[j] = (2 * 105.0 * (p[1] - p[0] * p[0])); j += 1


# lmdemo.c:70-96
def modros(p, x, m, n):
    for i in range(0, n, 3):
        x[i] = 10 * (p[1] - p[0] * p[0])
        x[i + 1] = 1.0 - p[0]
        x[i + 2] = 1e2


def jacmodros(p, jac, m, n):
    j = 0
    for i in range(0, n, 3):
        for v in (-20.0 * p[0], 10.0, -1.0, 0.0, 0.0, 0.0):
            jac[j] = v; j += 1


# lmdemo.c:99-120
def powell(p, x, m, n):
    for i in range(0, n, 2):
        x[i] = p[0]
        x[i + 1] = 10.0 * p[0] / (p[0] + 0.1) + 2 * p[1] * p[1]


def jacpowell(p, jac, m, n):
    j = 0
    for i in range(0, n, 2):
        for v in (1.0, 0.0, 1.0 / ((p[0] + 0.1) * (p[0] + 0.1)), 4.0 * p[1]):
            jac[j] = v; j += 1


# lmdemo.c:123-135
def wood(p, x, m, n):
    for i in range(0, n, 6):
        x[i] = 10.0 * (p[1] - p[0] * p[0])
        x[i + 1] = 1.0 - p[0]
        x[i + 2] = math.sqrt(90.0) * (p[3] - p[2] * p[2])
        x[i + 3] = 1.0 - p[2]
        x[i + 4] = math.sqrt(10.0) * (p[1] + p[3] - 2.0)
        x[i + 5] = (p[1] - p[3]) / math.sqrt(10.0)


# lmdemo.c:138-162
def meyer(p, x, m, n):
    for i in range(n):
        ui = 0.45 + 0.05 * i
        x[i] = p[0] * math.exp(10.0 * p[1] / (ui + p[2]) - 13.0)


def jacmeyer(p, jac, m, n):
    j = 0
    for i in range(n):
        ui = 0.45 + 0.05 * i
        tmp = math.exp(10.0 * p[1] / (ui + p[2]) - 13.0)
        jac[j] = tmp; j += 1
        jac[j] = 10.0 * p[0] * tmp / (ui + p[2]); j += 1
        jac[j] = -10.0 * p[0] * p[1] * tmp / ((ui + p[2]) * (ui + p[2])); j += 1


MEYER_X = [34.780, 28.610, 23.650, 19.630, 16.370, 13.720, 11.540, 9.744,
           8.261, 7.030, 6.005, 5.147, 4.427, 3.820, 3.307, 2.872]


# lmdemo.c:165-192
def osborne(p, x, m, n):
    for i in range(n):
        t = 10 * i
        x[i] = p[0] + p[1] * math.exp(-p[3] * t) + p[2] * math.exp(-p[4] * t)


def jacosborne(p, jac, m, n):
    j = 0
    for i in range(n):
        t = 10 * i
        tmp1 = math.exp(-p[3] * t)
        tmp2 = math.exp(-p[4] * t)
        for v in (1.0, tmp1, tmp2, -p[1] * t * tmp1, -p[2] * t * tmp2):
            jac[j] = v; j += 1


OSBORNE_X = [8.44E-1, 9.08E-1, 9.32E-1, 9.36E-1, 9.25E-1, 9.08E-1, 8.81E-1,
             8.5E-1, 8.18E-1, 7.84E-1, 7.51E-1, 7.18E-1, 6.85E-1, 6.58E-1,
             6.28E-1, 6.03E-1, 5.8E-1, 5.58E-1, 5.38E-1, 5.22E-1, 5.06E-1,
             4.9E-1, 4.78E-1, 4.67E-1, 4.57E-1, 4.48E-1, 4.38E-1, 4.31E-1,
             4.24E-1, 4.2E-1, 4.14E-1, 4.11E-1, 4.06E-1]

M_PI = 3.14159265358979323846


# lmdemo.c:200-233
def helval(p, x, m, n):
    if p[0] < 0.0:
        theta = math.atan(p[1] / p[0]) / (2.0 * M_PI) + 0.5
    elif 0.0 < p[0]:
        theta = math.atan(p[1] / p[0]) / (2.0 * M_PI)
    else:
        theta = 0.25 if p[1] >= 0 else -0.25
    x[0] = 10.0 * (p[2] - 10.0 * theta)
    x[1] = 10.0 * (math.sqrt(p[0] * p[0] + p[1] * p[1]) - 1.0)
    x[2] = p[2]


def jachelval(p, jac, m, n):
    tmp = p[0] * p[0] + p[1] * p[1]
    vals = (50.0 * p[1] / (M_PI * tmp), -50.0 * p[0] / (M_PI * tmp), 10.0,
            10.0 * p[0] / math.sqrt(tmp), 10.0 * p[1] / math.sqrt(tmp), 0.0,
            0.0, 0.0, 1.0)
    for i, v in enumerate(vals):
        jac[i] = v


# lmdemo.c:382-397
def hs01(p, x, m, n):
    t = p[0] * p[0]
    x[0] = 10.0 * (p[1] - t)
    x[1] = 1.0 - p[0]


def jachs01(p, jac, m, n):
    for i, v in enumerate((-20.0 * p[0], 10.0, -1.0, 0.0)):
        jac[i] = v


# lmdemo.c:407-421
def hs21(p, x, m, n):
    x[0] = p[0] / 10.0
    x[1] = p[1]


def jachs21(p, jac, m, n):
    for i, v in enumerate((0.1, 0.0, 0.0, 1.0)):
        jac[i] = v


def _sqrt(v):
    # C sqrt of a negative gives NaN instead of raising
    return math.sqrt(v) if v >= 0 else float("nan")


# lmdemo.c:428-460
def hatfldb(p, x, m, n):
    x[0] = p[0] - 1.0
    for i in range(1, m):
        x[i] = p[i - 1] - _sqrt(p[i])


def _div(a, b):
    if b == 0.0:
        return math.copysign(float("inf"), a) if a != 0 else float("nan")
    return a / b


def jachatfldb(p, jac, m, n):
    vals = (1.0, 0.0, 0.0, 0.0,
            1.0, _div(-0.5, _sqrt(p[1])), 0.0, 0.0,
            0.0, 1.0, _div(-0.5, _sqrt(p[2])), 0.0,
            0.0, 0.0, 1.0, _div(-0.5, _sqrt(p[3])))
    for i, v in enumerate(vals):
        jac[i] = v


# lmdemo.c:467-501
def hatfldc(p, x, m, n):
    x[0] = p[0] - 1.0
    for i in range(1, m - 1):
        x[i] = p[i - 1] - _sqrt(p[i])
    x[m - 1] = p[m - 1] - 1.0


def jachatfldc(p, jac, m, n):
    vals = (1.0, 0.0, 0.0, 0.0,
            1.0, _div(-0.5, _sqrt(p[1])), 0.0, 0.0,
            0.0, 1.0, _div(-0.5, _sqrt(p[2])), 0.0,
            0.0, 0.0, 0.0, 1.0)
    for i, v in enumerate(vals):
        jac[i] = v


# lmdemo.c:677-737
_R, _R5, _R6, _R7, _R8, _R9, _R10 = 10, 0.193, 4.10622 * 1e-4, 5.45177 * 1e-4, 4.4975 * 1e-7, 3.40735 * 1e-5, 9.615 * 1e-7


def combust(p, x, m, n):
    R, R5, R6, R7, R8, R9, R10 = _R, _R5, _R6, _R7, _R8, _R9, _R10
    x[0] = p[0] * p[1] + p[0] - 3 * p[4]
    x[1] = 2 * p[0] * p[1] + p[0] + 3 * R10 * p[1] * p[1] + p[1] * p[2] * p[2] + R7 * p[1] * p[2] + R9 * p[1] * p[3] + R8 * p[1] - R * p[4]
    x[2] = 2 * p[1] * p[2] * p[2] + R7 * p[1] * p[2] + 2 * R5 * p[2] * p[2] + R6 * p[2] - 8 * p[4]
    x[3] = R9 * p[1] * p[3] + 2 * p[3] * p[3] - 4 * R * p[4]
    x[4] = p[0] * p[1] + p[0] + R10 * p[1] * p[1] + p[1] * p[2] * p[2] + R7 * p[1] * p[2] + R9 * p[1] * p[3] + R8 * p[1] + R5 * p[2] * p[2] + R6 * p[2] + p[3] * p[3] - 1.0


def jaccombust(p, jac, m, n):
    R, R5, R6, R7, R8, R9, R10 = _R, _R5, _R6, _R7, _R8, _R9, _R10
    for j in range(m * n):
        jac[j] = 0.0
    j = 0
    jac[j] = p[1] + 1
    jac[j + 1] = p[0]
    jac[j + 4] = -3
    j += m
    jac[j] = 2 * p[1] + 1
    jac[j + 1] = 2 * p[0] + 6 * R10 * p[1] + p[2] * p[2] + R7 * p[2] + R9 * p[3] + R8
    jac[j + 2] = 2 * p[1] * p[2] + R7 * p[1]
    jac[j + 3] = R9 * p[1]
    jac[j + 4] = -R
    j += m
    jac[j + 1] = 2 * p[2] * p[2] + R7 * p[2]
    jac[j + 2] = 4 * p[1] * p[2] + R7 * p[1] + 4 * R5 * p[2] + R6
    jac[j + 4] = -8
    j += m
    jac[j + 1] = R9 * p[3]
    jac[j + 3] = R9 * p[1] + 4 * p[3]
    jac[j + 4] = -4 * R
    j += m
    jac[j] = p[1] + 1
    jac[j + 1] = p[0] + 2 * R10 * p[1] + p[2] * p[2] + R7 * p[2] + R9 * p[3] + R8
    jac[j + 2] = 2 * p[1] * p[2] + R7 * p[1] + 2 * R5 * p[2] + R6
    jac[j + 3] = R9 * p[1] + 2 * p[3]


def z(n):
    return np.zeros(n)


# (id, driver, func, jac, m, n, p0, x, lb, ub, itmax, expected solution (printed %.7g), expected info[5..9])
# lmdemo.c:859-1111; expectations from SURVEY.md section 4 (reference run in this container).
PROBLEMS = [
    dict(id=0, name="rosenbrock", driver="der", f=ros, j=jacros, m=2, n=2, p0=[-1.2, 1.0], x=z(2), itmax=1000,
         sol=None, info=None),
    dict(id=1, name="modros", driver="der", f=modros, j=jacmodros, m=2, n=3, p0=[-1.2, 1.0], x=z(3), itmax=1000,
         sol=[0.9999992, 0.9999984], info=[14, 2, 25, 14, 25]),
    dict(id=2, name="powell", driver="der", f=powell, j=jacpowell, m=2, n=2, p0=[3.0, 1.0], x=z(2), itmax=1000,
         sol=None, info=None),
    dict(id=3, name="wood", driver="dif", f=wood, j=None, m=4, n=6, p0=[-3.0, -1.0, -3.0, -1.0], x=z(6),
         itmax=1000, sol=[1, 1, 1, 1], info=[113, 6, 158, 11, 113]),
    dict(id=4, name="meyer", driver="dif", f=meyer, j=None, m=3, n=16, p0=[8.85, 4.0, 2.5],
         x=np.array(MEYER_X), itmax=1000, sol=[2.481778, 6.181346, 3.502236], info=[209, 2, 273, 21, 210],
         covar_row0=[0.00483514, -0.00162445, -0.000548114], info1=8.79459e-05),
    dict(id=5, name="osborne", driver="der", f=osborne, j=jacosborne, m=5, n=33,
         p0=[0.5, 1.5, -1.0, 1.0E-2, 2.0E-2], x=np.array(OSBORNE_X), itmax=1000,
         sol=[0.3754101, 1.935847, -1.464687, 0.01286753, 0.0221227], info=[34, 2, 45, 34, 45]),
    dict(id=6, name="helval", driver="der", f=helval, j=jachelval, m=3, n=3, p0=[-1.0, 0.0, 0.0], x=z(3),
         itmax=1000, sol=None, info=None),
    dict(id=11, name="hs01", driver="bc_der", f=hs01, j=jachs01, m=2, n=2, p0=[-2.0, 1.0], x=z(2),
         lb=[-DBL_MAX, -1.5], ub=[DBL_MAX, DBL_MAX], itmax=1000, sol=[1, 1], info=[14, 6, 23, 14, 14]),
    dict(id=12, name="hs21mod", driver="bc_der", f=hs21, j=jachs21, m=2, n=2, p0=[-1.0, -1.0], x=z(2),
         lb=[2.0, -50.0], ub=[50.0, 50.0], itmax=1000, sol=[2, -4.688186e-19], info=[5, 1, 10, 6, 5]),
    dict(id=13, name="hatfldb", driver="bc_der", f=hatfldb, j=jachatfldb, m=4, n=4, p0=[0.1] * 4, x=z(4),
         lb=[0.0] * 4, ub=[DBL_MAX, 0.8, DBL_MAX, DBL_MAX], itmax=1000,
         sol=[0.9472136, 0.8, 0.64, 0.4096], info=[939, 2, 3186, 939, 939]),
    dict(id=14, name="hatfldc", driver="bc_der", f=hatfldc, j=jachatfldc, m=4, n=4, p0=[0.9] * 4, x=z(4),
         lb=[0.0] * 4, ub=[10.0] * 4, itmax=1000, sol=[1, 1, 1, 1], info=[4, 6, 5, 4, 4]),
    dict(id=15, name="combust", driver="bc_der", f=combust, j=jaccombust, m=5, n=5, p0=[0.0001] * 5, x=z(5),
         lb=[0.0001] * 5, ub=[100.0] * 5, itmax=5000,
         sol=[0.00343023, 31.3265, 0.0683504, 0.859529, 0.03696244], info=[68, 6, 87, 68, 68]),
]


def callbacks(prob):
    f = _wrap(prob["f"])
    j = _wrap(prob["j"]) if prob["j"] is not None else None
    return f, j
