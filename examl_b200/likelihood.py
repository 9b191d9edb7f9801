"""Host-side mirror of the three axml.h entry points on top of the C-ABI engine.

`Likelihood` plays the role of ExaML's `tree` for this path: it owns the tree, the partition
masks and the Newton-Raphson bookkeeping, builds traversal descriptors on the host and hands
them to the engine -- the same split the C drop-in (examl_b200/csrc/axml_backend.c) makes.
Names, argument meaning and side effects follow the reference:

  evaluateGeneric(p, fullTraversal)                 evaluateGenericSpecial.c:897-1003
  newviewGeneric(p, masked)                         newviewGenericSpecial.c:1523-1593
  makenewzGeneric(p, q, z0, maxiter, mask)          makenewzGenericSpecial.c:1355-1398
  topLevelMakenewz (the NR state machine)           makenewzGenericSpecial.c:1133-1349
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from .phylo import ZMAX, ZMIN, Entry, Node, Tree


class Likelihood:
    def __init__(self, tree: Tree, engine, num_partitions: int, device_newton_raphson: bool = False):
        self.device_newton_raphson = device_newton_raphson   # run topLevelMakenewz on the GPU (single rank)
        self.tr = tree
        self.engine = engine
        self.NumberOfModels = num_partitions
        self.numBranches = tree.numBranches
        self.executeModel = [True] * num_partitions
        self.partitionConverged = [False] * max(num_partitions, self.numBranches)
        self.perPartitionLH = np.zeros(num_partitions)
        self.likelihood = 0.0
        self.curvatOK = [True] * self.numBranches
        self.nr_iterations = 0          # execCore calls of the last makenewzGeneric (diagnostics)

    # ---- evaluateGeneric ------------------------------------------------------------------------
    def evaluateGeneric(self, p: Node, fullTraversal: bool) -> float:
        if fullTraversal:
            ti = self.tr.full_traversal(p)
        else:
            ti = self.tr.partial_traversal(p)
        self.engine.evaluateIterative(ti, self.executeModel, out=self.perPartitionLH)
        self.likelihood = float(sum(self.perPartitionLH[m] for m in range(self.NumberOfModels)))
        return self.likelihood

    # ---- newviewGeneric ---------------------------------------------------------------------------
    def newviewGeneric(self, p: Node, masked: bool = False) -> None:
        if self.tr.isTip(p.number):
            return
        ti = self.tr.newview_traversal(p)
        if masked:
            for m in range(self.NumberOfModels):
                self.executeModel[m] = not self.partitionConverged[m]
        if ti:
            self.engine.newviewIterative(ti, self.executeModel)
        if masked:
            self.executeModel = [True] * self.NumberOfModels

    # ---- makenewzGeneric ----------------------------------------------------------------------------
    def makenewzGeneric(self, p: Node, q: Node, z0: Sequence[float], maxiter: int, mask: bool = False) -> List[float]:
        nb = self.numBranches
        for i in range(nb):
            if mask:
                self.executeModel[i] = not self.partitionConverged[i]
            else:
                assert self.executeModel[i]
        ti = [Entry(-1, p.number, q.number, 0, list(z0[:nb]), [0.0] * nb)]
        if not self.tr.isTip(p.number) and not p.x:
            self.tr.computeTraversalInfo(p, ti, True)
        if not self.tr.isTip(q.number) and not q.x:
            self.tr.computeTraversalInfo(q, ti, True)
        if self.device_newton_raphson:
            z, cok, self.nr_iterations = self.engine.makenewz(ti, list(z0[:nb]), maxiter, self.executeModel)
            result, self.curvatOK = [float(v) for v in z], [bool(v) for v in cok]
            self.executeModel = [True] * self.NumberOfModels
        else:
            result = self._topLevelMakenewz(ti, list(z0[:nb]), maxiter)
        for i in range(nb):
            self.executeModel[i] = True
        return result

    def _topLevelMakenewz(self, ti: List[Entry], z0: List[float], _maxiter: int) -> List[float]:
        nb = self.numBranches
        z = list(z0)
        zprev = [0.0] * nb
        zstep = [0.0] * nb
        maxiter = [_maxiter] * nb
        outerConverged = [False] * nb
        coreLZ = [0.0] * nb
        self.curvatOK = [True] * nb
        firstIteration = True
        self.nr_iterations = 0
        while True:
            for i in range(nb):
                if not outerConverged[i] and self.curvatOK[i]:
                    self.curvatOK[i] = False
                    zprev[i] = z[i]
                    zstep[i] = (1.0 - ZMAX) * z[i] + ZMIN
            for i in range(nb):
                if not outerConverged[i] and not self.curvatOK[i]:
                    z[i] = min(max(z[i], ZMIN), ZMAX)
                    coreLZ[i] = math.log(z[i])
            if nb > 1:
                assert nb == self.NumberOfModels
                for m in range(self.NumberOfModels):
                    if self.executeModel[m]:
                        self.executeModel[m] = not self.curvatOK[m]
            else:
                for m in range(self.NumberOfModels):
                    self.executeModel[m] = not self.curvatOK[0]
            if firstIteration:
                self.engine.makenewzIterative(ti, self.executeModel)
                firstIteration = False
            dlnLdlz, d2lnLdlz2 = self.engine.execCore(coreLZ, self.executeModel)
            self.nr_iterations += 1
            for i in range(nb):
                if not outerConverged[i] and not self.curvatOK[i]:
                    if d2lnLdlz2[i] >= 0.0 and z[i] < ZMAX:
                        zprev[i] = z[i] = 0.37 * z[i] + 0.63      # bad curvature: shorten the branch
                    else:
                        self.curvatOK[i] = True
            for i in range(nb):
                if self.curvatOK[i] and not outerConverged[i]:
                    if d2lnLdlz2[i] < 0.0:
                        tantmp = -dlnLdlz[i] / d2lnLdlz2[i]
                        if tantmp < 100:
                            z[i] *= math.exp(tantmp)
                            if z[i] < ZMIN:
                                z[i] = ZMIN
                            if z[i] > 0.25 * zprev[i] + 0.75:
                                z[i] = 0.25 * zprev[i] + 0.75
                        else:
                            z[i] = 0.25 * zprev[i] + 0.75
                    if z[i] > ZMAX:
                        z[i] = ZMAX
                    maxiter[i] -= 1
                    if abs(z[i] - zprev[i]) > zstep[i]:
                        if maxiter[i] < -20:
                            z[i] = z0[i]
                            outerConverged[i] = True
                        else:
                            outerConverged[i] = False
                    else:
                        outerConverged[i] = True
            if all(outerConverged):
                break
        self.executeModel = [True] * self.NumberOfModels
        return z
