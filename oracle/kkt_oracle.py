"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the SLP iteration's KKT / merit metrics:
KT_residuals (src/opt/algorithms/common.jl:35-44), norm_complementarity (:51-69), norm_violations (:76-98),
compute_phi at alpha = 0 and compute_derivative outside feasibility restoration (src/opt/algorithms/slp.jl:108-131,
138-162).  numpy, written statement by statement after the Julia.  Parity unpinned against a Julia run."""
import numpy as np


def KT_residuals(df, lam, mult_x_U, mult_x_L, Jac):
    KT_res = np.linalg.norm(df - Jac.T @ lam - mult_x_U - mult_x_L)        # :38
    scalar = max(1.0, np.linalg.norm(df))                                  # :39
    for i in range(Jac.shape[0]):                                          # :40-42
        scalar = max(scalar, abs(lam[i]) * np.linalg.norm(Jac[i, :]))
    return KT_res / scalar


def norm_complementarity(E, g_L, g_U, lam, p=np.inf):
    m = len(E)
    compl = np.zeros(m)
    denom = 0.0
    for i in range(m):                                                     # :58-67
        if g_L[i] == g_U[i]:
            compl[i] = 0.0
        else:
            with np.errstate(invalid="ignore"):
                c = min(E[i] - g_L[i], g_U[i] - E[i]) * lam[i]
            compl[i] = 0.0 if np.isnan(c) else c
            denom += lam[i] ** 2
    return np.linalg.norm(compl, p) / (1 + np.sqrt(denom))                 # :68


def norm_violations(E, g_L, g_U, x, x_L, x_U, p=1):
    m, n = len(E), len(x)
    viol = np.zeros(m + n)
    for i in range(m):                                                     # :84-90
        if E[i] > g_U[i]:
            viol[i] = E[i] - g_U[i]
        elif E[i] < g_L[i]:
            viol[i] = g_L[i] - E[i]
    for j in range(n):                                                     # :91-97
        if x[j] > x_U[j]:
            viol[m + j] = x[j] - x_U[j]
        elif x[j] < x_L[j]:
            viol[m + j] = x_L[j] - x[j]
    return np.linalg.norm(viol, p)


def compute_phi0(f, E, g_L, g_U, nu):
    phi = f                                                                # slp.jl:121
    for i in range(len(E)):                                                # :122-128
        if E[i] > g_U[i]:
            phi += nu[i] * (E[i] - g_U[i])
        elif E[i] < g_L[i]:
            phi += nu[i] * (g_L[i] - E[i])
    return phi


def compute_derivative(df, p, E, g_L, g_U, nu):
    D = float(df @ p)                                                      # slp.jl:153
    for i in range(len(E)):                                                # :154-160
        if E[i] > g_U[i]:
            D -= nu[i] * (E[i] - g_U[i])
        elif E[i] < g_L[i]:
            D -= nu[i] * (g_L[i] - E[i])
    return D
