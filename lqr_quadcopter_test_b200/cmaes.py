"""(mu/mu_w, lambda)-CMA-ES ask/tell on the host.

The reference drives the third-party ``cma`` package (``cma.CMAEvolutionStrategy(x0, sigma0, opts)``,
controllers/tuning.py:1313-1428; the package is an optional dependency, ``pyproject.toml`` extra, and is
not installed in this image).  This module restates the published algorithm the package implements
(N. Hansen, "The CMA Evolution Strategy: A Tutorial", 2016: default population size, recombination
weights, step-size path with the h_sigma stall, rank-one and rank-mu covariance updates) for the three
things the tuner uses: ``ask()``, ``tell(solutions, fitnesses)`` and ``stop()``.  Box bounds are handled
by repair + quadratic penalty (evaluate the point clipped into the box, penalise the distance), not by
the package's coordinate transform, and the random stream is numpy's ``default_rng(seed)`` — so the
sequence of candidates is NOT the package's; **parity with ``cma`` is unpinned** (it cannot be run
here).  What stays pinned is everything around it: parameter specification, bounds, x0, sigma0
(tuning.py:1156-1262, 1337-1342) and the evaluation of each candidate, which is the batched GPU path.
The whole population of one generation is evaluated in one DARE + rollout launch.
"""
from __future__ import annotations

import numpy as np


class CMAEvolutionStrategy:
    def __init__(self, x0, sigma0: float, lower, upper, seed: int = 0, popsize: int | None = None,
                 tolfun: float = 1e-11, tolx: float = 1e-12, maxfevals: int | None = None):
        self.mean = np.asarray(x0, dtype=np.float64).copy()
        n = self.n = self.mean.size
        self.sigma = float(sigma0)
        self.lower, self.upper = np.asarray(lower, dtype=np.float64), np.asarray(upper, dtype=np.float64)
        self.rng = np.random.default_rng(seed)
        self.lam = int(popsize) if popsize is not None else 4 + int(3 * np.log(n))
        self.mu = self.lam // 2
        w = np.log((self.lam + 1) / 2.0) - np.log(np.arange(1, self.mu + 1))
        self.weights = w / w.sum()
        self.mueff = 1.0 / np.sum(self.weights ** 2)
        self.cc = (4 + self.mueff / n) / (n + 4 + 2 * self.mueff / n)
        self.cs = (self.mueff + 2) / (n + self.mueff + 5)
        self.c1 = 2 / ((n + 1.3) ** 2 + self.mueff)
        self.cmu = min(1 - self.c1, 2 * (self.mueff - 2 + 1 / self.mueff) / ((n + 2) ** 2 + self.mueff))
        self.damps = 1 + 2 * max(0.0, np.sqrt((self.mueff - 1) / (n + 1)) - 1) + self.cs
        self.chiN = np.sqrt(n) * (1 - 1 / (4 * n) + 1 / (21 * n * n))
        # coordinate scales: the search box is anisotropic (weights span 1e-4 .. 25), so the initial
        # covariance is diag(scale^2) with scale_i proportional to the width of the box in coordinate i
        width = np.where(self.upper > self.lower, self.upper - self.lower, 1.0)
        self.scale = width / np.exp(np.mean(np.log(width)))
        self.pc, self.ps = np.zeros(n), np.zeros(n)
        self.C = np.eye(n)
        self.B, self.D = np.eye(n), np.ones(n)
        self.counteval = 0
        self.generation = 0
        self.tolfun, self.tolx, self.maxfevals = tolfun, tolx, maxfevals
        self._fit_hist: list[float] = []
        self._last_z = None
        self.best_x, self.best_f = None, np.inf

    def repair(self, x: np.ndarray) -> np.ndarray:
        return np.minimum(np.maximum(x, self.lower), self.upper)

    def ask(self) -> list[np.ndarray]:
        z = self.rng.standard_normal((self.lam, self.n))
        y = (z * self.D) @ self.B.T
        self._last_y = y
        return [self.mean + self.sigma * self.scale * yi for yi in y]

    def penalised(self, x: np.ndarray, f: float) -> float:
        """Fitness to pass to tell() for a candidate that was evaluated at repair(x)."""
        d = (x - self.repair(x)) / (self.scale * max(self.sigma, 1e-300))
        return float(f + 1e-3 * float(d @ d) * (1.0 + abs(f)))

    def tell(self, solutions, fitnesses) -> None:
        X = np.asarray(solutions, dtype=np.float64)
        f = np.asarray(fitnesses, dtype=np.float64)
        k = len(f)
        self.counteval += k
        order = np.argsort(f, kind="stable")
        if f[order[0]] < self.best_f:
            self.best_f, self.best_x = float(f[order[0]]), self.repair(X[order[0]]).copy()
        mu = min(self.mu, k)
        w = self.weights[:mu] / self.weights[:mu].sum()
        Y = (X[order[:mu]] - self.mean) / (self.sigma * self.scale)
        yw = w @ Y
        self.mean = self.mean + self.sigma * self.scale * yw
        n = self.n
        invsqrtC_yw = self.B @ ((self.B.T @ yw) / self.D)
        self.ps = (1 - self.cs) * self.ps + np.sqrt(self.cs * (2 - self.cs) * self.mueff) * invsqrtC_yw
        self.generation += 1
        hsig = (np.linalg.norm(self.ps) / np.sqrt(1 - (1 - self.cs) ** (2 * self.generation)) / self.chiN) < 1.4 + 2 / (n + 1)
        self.pc = (1 - self.cc) * self.pc + hsig * np.sqrt(self.cc * (2 - self.cc) * self.mueff) * yw
        rank_mu = (Y.T * w) @ Y
        self.C = ((1 - self.c1 - self.cmu) * self.C
                  + self.c1 * (np.outer(self.pc, self.pc) + (1 - hsig) * self.cc * (2 - self.cc) * self.C)
                  + self.cmu * rank_mu)
        self.sigma *= float(np.exp((self.cs / self.damps) * (np.linalg.norm(self.ps) / self.chiN - 1)))
        self.C = np.triu(self.C) + np.triu(self.C, 1).T
        vals, vecs = np.linalg.eigh(self.C)
        self.D = np.sqrt(np.maximum(vals, 1e-30))
        self.B = vecs
        self._fit_hist.append(float(f[order[0]]))

    def stop(self) -> dict:
        reasons = {}
        if self.maxfevals is not None and self.counteval >= self.maxfevals:
            reasons["maxfevals"] = self.maxfevals
        if self.generation > 0 and self.sigma * float(np.max(self.D * self.scale)) < self.tolx:
            reasons["tolx"] = self.tolx
        h = self._fit_hist[-(10 + int(30 * self.n / self.lam)):]
        if len(h) >= 10 + int(30 * self.n / self.lam) and max(h) - min(h) < self.tolfun:
            reasons["tolfun"] = self.tolfun
        return reasons
