"""Legendre-Gauss-Lobatto basis and DGSEM operators (host side, built once, uploaded at create).

Restates, in NumPy, what the reference computes in
``src/solvers/dgsem/basis_lobatto_legendre.jl``:
  * gauss_lobatto_nodes_weights   :570-639  (Kopriva alg. 25, Newton iteration)
  * calc_q_and_l                  :642-664  (Kopriva alg. 24)
  * barycentric_weights           :500-514
  * polynomial_derivative_matrix  :427-440  (Kopriva alg. 37)
  * calc_Dhat                     :399-409  (= -M^-1 D^T M)
  * calc_Dsplit                   :414-423  (= 2D - M^-1 B, zero diagonal)
  * polynomial_interpolation_matrix :444-484 (Kopriva alg. 32)
  * SolutionAnalyzer              :274-290  (analysis_polydeg = 2*polydeg)
All matrices are plain float64, indexed [row, col] like the reference's Matrix{Float64}.
"""
from __future__ import annotations

import math
import numpy as np


def calc_q_and_l(N: int, x: float):
    L_Nm2, L_Nm1 = 1.0, x
    Lder_Nm2, Lder_Nm1 = 0.0, 1.0
    L = x
    for i in range(2, N + 1):
        L = ((2 * i - 1) * x * L_Nm1 - (i - 1) * L_Nm2) / i
        Lder = Lder_Nm2 + (2 * i - 1) * L_Nm1
        L_Nm2, L_Nm1 = L_Nm1, L
        Lder_Nm2, Lder_Nm1 = Lder_Nm1, Lder
    q = (2 * N + 1) / (N + 1) * (x * L - L_Nm2)
    qder = (2 * N + 1) * L
    return q, qder, L


def gauss_lobatto_nodes_weights(n_nodes: int):
    n_iterations = 20
    tolerance = 2 * np.finfo(np.float64).eps
    nodes = np.zeros(n_nodes)
    weights = np.zeros(n_nodes)
    if n_nodes == 1:
        nodes[0] = 0.0
        weights[0] = 2.0
        return nodes, weights
    N = n_nodes - 1
    nodes[0], nodes[-1] = -1.0, 1.0
    weights[0] = 2.0 / (N * (N + 1))
    weights[-1] = weights[0]
    if N > 1:
        cont1 = math.pi / N
        cont2 = 3 / (8 * N * math.pi)
        for i in range(1, (N + 1) // 2):
            x = -math.cos(cont1 * (i + 0.25) - cont2 / (i + 0.25))
            for _ in range(n_iterations + 1):
                q, qder, _L = calc_q_and_l(N, x)
                dx = -q / qder
                x += dx
                if abs(dx) < tolerance * abs(x):
                    break
            _, _, L = calc_q_and_l(N, x)
            nodes[i] = x
            weights[i] = weights[0] / L**2
            nodes[N - i] = -x
            weights[N - i] = weights[i]
    if n_nodes % 2 == 1:
        _, _, L = calc_q_and_l(N, 0.0)
        nodes[N // 2] = 0.0
        weights[N // 2] = weights[0] / L**2
    return nodes, weights


def legendre_polynomial_and_derivative(N: int, x: float):
    # basis_lobatto_legendre.jl:673-699 (Kopriva alg. 22), normalised
    if N == 0:
        poly, deriv = 1.0, 0.0
    elif N == 1:
        poly, deriv = x, 1.0
    else:
        p2, p1 = 1.0, x
        d2, d1 = 0.0, 1.0
        poly = deriv = 0.0
        for i in range(2, N + 1):
            poly = ((2 * i - 1) * x * p1 - (i - 1) * p2) / i
            deriv = d2 + (2 * i - 1) * p1
            p2, p1 = p1, poly
            d2, d1 = d1, deriv
    poly *= math.sqrt(N + 0.5)
    deriv *= math.sqrt(N + 0.5)
    return poly, deriv


def gauss_nodes_weights(n_nodes: int):
    # basis_lobatto_legendre.jl:702-770 (Kopriva alg. 23)
    n_iterations = 20
    tolerance = 2 * np.finfo(np.float64).eps
    nodes = np.ones(n_nodes)
    weights = np.zeros(n_nodes)
    N = n_nodes - 1
    if N == 0:
        nodes[0] = 0.0
        weights[0] = 2.0
    elif N == 1:
        nodes[0] = -math.sqrt(1 / 3)
        nodes[1] = -nodes[0]
        weights[:] = 1.0
    else:
        for i in range(0, (N + 1) // 2):
            x = -math.cos(math.pi / (2 * N + 2) * (2 * i + 1))
            for _ in range(n_iterations + 1):
                poly, deriv = legendre_polynomial_and_derivative(N + 1, x)
                dx = -poly / deriv
                x += dx
                if abs(dx) < tolerance * abs(x):
                    break
            poly, deriv = legendre_polynomial_and_derivative(N + 1, x)
            nodes[i] = x
            weights[i] = (2 * N + 3) / ((1 - x**2) * deriv**2)
            nodes[N - i] = -x
            weights[N - i] = weights[i]
        if n_nodes % 2 == 1:
            poly, deriv = legendre_polynomial_and_derivative(N + 1, 0.0)
            nodes[N // 2] = 0.0
            weights[N // 2] = (2 * N + 3) / deriv**2
    return nodes, weights


def barycentric_weights(nodes):
    n = len(nodes)
    w = np.ones(n)
    for j in range(1, n):
        for k in range(j):
            w[k] *= nodes[k] - nodes[j]
            w[j] *= nodes[j] - nodes[k]
    return 1.0 / w


def polynomial_derivative_matrix(nodes):
    n = len(nodes)
    D = np.zeros((n, n))
    wbary = barycentric_weights(nodes)
    for i in range(n):
        for j in range(n):
            if j != i:
                D[i, j] = (wbary[j] / wbary[i]) * 1 / (nodes[i] - nodes[j])
                D[i, i] -= D[i, j]
    return D


def calc_Dhat(D, weights):
    n = len(weights)
    Dhat = D.T.copy()
    for nn in range(n):
        for j in range(n):
            Dhat[j, nn] *= -weights[nn] / weights[j]
    return Dhat


def calc_Dsplit(D, weights):
    Ds = 2.0 * D
    Ds[0, 0] += 1 / weights[0]
    Ds[-1, -1] -= 1 / weights[-1]
    return Ds


def _isapprox(a, b):
    return abs(a - b) <= math.sqrt(np.finfo(np.float64).eps) * max(abs(a), abs(b))


def polynomial_interpolation_matrix(nodes_in, nodes_out, baryweights_in=None):
    if baryweights_in is None:
        baryweights_in = barycentric_weights(nodes_in)
    V = np.zeros((len(nodes_out), len(nodes_in)))
    for k, xo in enumerate(nodes_out):
        match = False
        for j, xi in enumerate(nodes_in):
            if _isapprox(xo, xi):
                match = True
                V[k, j] = 1.0
        if not match:
            s = 0.0
            for j, xi in enumerate(nodes_in):
                t = baryweights_in[j] / (xo - xi)
                V[k, j] = t
                s += t
            V[k, :] /= s
    return V


def lagrange_interpolating_polynomials(x, nodes, wbary):
    # basis_lobatto_legendre.jl:523-554
    n = len(nodes)
    poly = np.zeros(n)
    for i in range(n):
        if _isapprox(x, nodes[i]):
            poly[i] = 1.0
            return poly
    for i in range(n):
        poly[i] = wbary[i] / (x - nodes[i])
    return poly / poly.sum()


def legendre_polynomial(N, x):
    """Normalised Legendre polynomial of degree N (legendre_polynomial_and_derivative,
    basis_lobatto_legendre.jl:677-708: three-term recurrence, scaled by sqrt(N + 1/2))."""
    if N == 0:
        poly = 1.0
    elif N == 1:
        poly = x
    else:
        poly_nm2, poly_nm1 = 1.0, x
        poly = 0.0
        for i in range(2, N + 1):
            poly = ((2 * i - 1) * x * poly_nm1 - (i - 1) * poly_nm2) / i
            poly_nm2, poly_nm1 = poly_nm1, poly
    return poly * math.sqrt(N + 0.5)


def vandermonde_legendre(nodes):
    """``vandermonde_legendre(nodes)`` (basis_lobatto_legendre.jl:711-727): nodal -> modal is its inverse."""
    n = len(nodes)
    V = np.array([[legendre_polynomial(m, float(nodes[i])) for m in range(n)] for i in range(n)])
    return V, np.linalg.inv(V)


class LobattoLegendreBasis:
    """Mirror of ``LobattoLegendreBasis`` (basis_lobatto_legendre.jl:17-86)."""

    def __init__(self, polydeg: int):
        n = polydeg + 1
        self.polydeg = polydeg
        self.nodes, self.weights = gauss_lobatto_nodes_weights(n)
        self.inverse_weights = 1.0 / self.weights
        self.derivative_matrix = polynomial_derivative_matrix(self.nodes)
        self.derivative_split = calc_Dsplit(self.derivative_matrix, self.weights)
        self.derivative_hat = calc_Dhat(self.derivative_matrix, self.weights)
        _, self.inverse_vandermonde_legendre = vandermonde_legendre(self.nodes)

    @property
    def nnodes(self):
        return self.polydeg + 1


class LobattoLegendreMortarL2:
    """``MortarL2(basis)`` (basis_lobatto_legendre.jl:159-206): interpolation from a large face to its
    lower/upper half (``calc_forward_*`` l2projection.jl:26-61) and the discrete L2 projection back, evaluated
    on Gauss nodes (``calc_reverse_*(n, Val(:gauss))`` l2projection.jl:67-115)."""

    def __init__(self, basis):
        n = basis.nnodes
        nodes = basis.nodes
        wbary = barycentric_weights(nodes)
        self.forward_upper = np.array([lagrange_interpolating_polynomials(0.5 * (x + 1), nodes, wbary) for x in nodes])
        self.forward_lower = np.array([lagrange_interpolating_polynomials(0.5 * (x - 1), nodes, wbary) for x in nodes])
        gn, gw = gauss_nodes_weights(n)
        gbary = barycentric_weights(gn)
        g2l = polynomial_interpolation_matrix(gn, nodes)
        l2g = polynomial_interpolation_matrix(nodes, gn)

        def reverse(shift):
            op = np.zeros((n, n))
            for j in range(n):
                poly = lagrange_interpolating_polynomials(0.5 * (gn[j] + shift), gn, gbary)
                for i in range(n):
                    op[i, j] = 0.5 * poly[i] * gw[j] / gw[i]
            return g2l @ op @ l2g

        self.reverse_upper = reverse(+1.0)
        self.reverse_lower = reverse(-1.0)


class SolutionAnalyzer:
    """Mirror of ``SolutionAnalyzer`` (basis_lobatto_legendre.jl:274-290)."""

    def __init__(self, basis: LobattoLegendreBasis, analysis_polydeg=None):
        if analysis_polydeg is None:
            analysis_polydeg = 2 * basis.polydeg
        self.nodes, self.weights = gauss_lobatto_nodes_weights(analysis_polydeg + 1)
        self.vandermonde = polynomial_interpolation_matrix(basis.nodes, self.nodes)
