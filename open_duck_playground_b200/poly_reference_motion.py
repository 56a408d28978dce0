"""Polynomial reference motion table (host side).

Mirrors ``PolyReferenceMotion.process`` (reference common/poly_reference_motion.py:74-146):
the pickle holds, per ``"dx_dy_dtheta"`` key, 40 polynomials of 16 coefficients
(lowest power first); the reference flips them (``:113``) and stacks a dense
``[ndx, ndy, ndtheta, 40, 16]`` tensor indexed by the sorted unique grid values.
The per-step evaluation (nearest grid cell + Horner, ``:148-168``) runs inside the
CUDA step kernel; this module only builds the table it reads.
"""
from __future__ import annotations

import pickle
from dataclasses import dataclass
from typing import List

import numpy as np


@dataclass
class PolyTable:
    dxs: List[float]
    dys: List[float]
    dthetas: List[float]
    dx_range: List[float]
    dy_range: List[float]
    dtheta_range: List[float]
    period: float
    fps: float
    nb_steps_in_period: int
    coef: np.ndarray  # float64 [ndx, ndy, ndth, 40, 16], highest power first (jp.polyval order)

    def save(self, path: str) -> None:
        np.savez_compressed(path, dxs=self.dxs, dys=self.dys, dthetas=self.dthetas, dx_range=self.dx_range,
                            dy_range=self.dy_range, dtheta_range=self.dtheta_range, period=self.period, fps=self.fps,
                            nb_steps_in_period=self.nb_steps_in_period, coef=self.coef)

    @staticmethod
    def load(path: str) -> "PolyTable":
        z = np.load(path)
        return PolyTable(dxs=z["dxs"].tolist(), dys=z["dys"].tolist(), dthetas=z["dthetas"].tolist(),
                         dx_range=z["dx_range"].tolist(), dy_range=z["dy_range"].tolist(),
                         dtheta_range=z["dtheta_range"].tolist(), period=float(z["period"]), fps=float(z["fps"]),
                         nb_steps_in_period=int(z["nb_steps_in_period"]), coef=z["coef"])

    @staticmethod
    def from_pickle(path: str) -> "PolyTable":
        data = pickle.load(open(path, "rb"))
        dx_range, dy_range, dth_range = [0.0, 0.0], [0.0, 0.0], [0.0, 0.0]  # ranges start at [0, 0] (:59-61)
        dxs, dys, dths = [], [], []
        period = fps = None
        cells = {}
        for name, entry in data.items():
            dx, dy, dth = (float(s) for s in name.split("_"))
            if period is None:
                period, fps = entry["period"], entry["fps"]
            for v, lst in ((dx, dxs), (dy, dys), (dth, dths)):
                if v not in lst:
                    lst.append(v)
            dx_range = [min(dx, dx_range[0]), max(dx, dx_range[1])]
            dy_range = [min(dy, dy_range[0]), max(dy, dy_range[1])]
            dth_range = [min(dth, dth_range[0]), max(dth, dth_range[1])]
            cells[(dx, dy, dth)] = np.array([np.flip(np.asarray(v, dtype=np.float64)) for v in entry["coefficients"].values()])
        dxs, dys, dths = sorted(dxs), sorted(dys), sorted(dths)
        coef = np.zeros((len(dxs), len(dys), len(dths), 40, 16))
        for ix, dx in enumerate(dxs):
            for iy, dy in enumerate(dys):
                for it, dth in enumerate(dths):
                    coef[ix, iy, it] = cells[(dx, dy, dth)]
        return PolyTable(dxs, dys, dths, dx_range, dy_range, dth_range, float(period), float(fps),
                         int(period * fps), coef)
