"""Host-side mirror of the reference's driver loop: ``Solver<LP_TYPE,VISITOR>::Solve`` and ``MpRoundingSolver``
(/root/reference/include/solver.hxx:230-340,381-400) and the ``StandardVisitor`` control law
(include/visitors/standard_visitor.hxx:48-199).  Option names and defaults are the reference's TCLAP options
(standard_visitor.hxx:32-44); the log lines are the reference's (``iteration = k, lower bound = ...``).

This is host logic only: all arithmetic happens behind ``LP`` on the device.
"""
from __future__ import annotations

import math
import time
from dataclasses import dataclass, field

EPS = 1e-8  # include/config.hxx:51 (REAL = double)
INDEX_MAX = 2 ** 64 - 1


@dataclass
class LpControl:  # include/config.hxx:95-105
    repam: str | None = None
    computePrimal: bool = False
    computeLowerBound: bool = False
    tighten: bool = False
    end: bool = False
    error: bool = False


@dataclass
class StandardVisitor:
    """include/visitors/standard_visitor.hxx:25-253 (tightening visitor: out of scope)."""
    maxIter: int = 1000
    maxMemory: int = INDEX_MAX
    timeout: int = INDEX_MAX
    primalComputationInterval: int = 5
    primalComputationStart: int = 1
    lowerBoundComputationInterval: int = 1
    minDualImprovement: float | None = None   # None == option not set (minDualImprovementArg_.isSet())
    minDualImprovementInterval: int = 10
    standardReparametrization: str = "anisotropic"
    roundingReparametrization: str = "damped_uniform"
    verbosity: int = 0
    lowerBound_: list = field(default_factory=list)
    log: list = field(default_factory=list)

    def begin(self, lp) -> LpControl:  # :48-106
        self.remainingIter_ = self.maxIter
        self.curIter_ = 0
        self.lowerBound_ = []
        self.beginTime_ = time.monotonic()
        return LpControl(repam=self.standardReparametrization, computePrimal=False, computeLowerBound=True)

    def _say(self, s):
        self.log.append(s)
        if self.verbosity >= 1:
            print(s)

    def visit(self, c: LpControl, lowerBound: float, primalBound: float) -> LpControl:  # :110-190
        self.lowerBound_.append(lowerBound)
        elapsed_ms = int((time.monotonic() - self.beginTime_) * 1000)
        if c.computePrimal or c.computeLowerBound:
            s = f"iteration = {self.curIter_}"
            if c.computeLowerBound:
                s += f", lower bound = {lowerBound:.10g}"
            if c.computePrimal:
                s += f", upper bound = {primalBound:.10g}"
            self._say(s + f", time elapsed = {elapsed_ms // 1000}.{(elapsed_ms % 1000) // 10}s")
        self.curIter_ += 1
        self.remainingIter_ -= 1
        ret = LpControl()
        if self.remainingIter_ == 0:
            self._say("One iteration remaining")
            ret.end = True
            return ret
        if primalBound <= lowerBound + EPS:
            self._say(f"Primal cost {primalBound:.10g} greater equal lower bound {lowerBound:.10g}")
            ret.end = True
            return ret
        if self.timeout != INDEX_MAX and elapsed_ms // 1000 >= self.timeout:
            self._say(f"Timeout reached after {elapsed_ms} seconds")
            self.remainingIter_ = min(1, self.remainingIter_)
        if (c.computeLowerBound and self.curIter_ >= self.minDualImprovementInterval
                and self.minDualImprovement is not None):
            # the reference indexes lowerBound_[size-1-interval]; at the first eligible visit that is index -1 (an
            # out-of-bounds read, standard_visitor.hxx:166-167) -- we skip that one visit instead of reading garbage
            k = len(self.lowerBound_) - 1 - self.minDualImprovementInterval
            prev = self.lowerBound_[k] if k >= 0 else -math.inf
            if self.minDualImprovement > 0 and lowerBound - prev < self.minDualImprovement:
                self._say(f"Dual improvement smaller than {self.minDualImprovement} after "
                          f"{self.minDualImprovementInterval} iterations, terminating optimization")
                self.remainingIter_ = min(1, self.remainingIter_)
        if self.remainingIter_ == 1:
            return LpControl(repam=self.roundingReparametrization, computePrimal=True, computeLowerBound=True)
        ret.repam = self.standardReparametrization
        if (self.curIter_ >= self.primalComputationStart
                and (self.curIter_ - self.primalComputationStart) % self.primalComputationInterval == 0):
            ret.computePrimal = True
            ret.repam = self.roundingReparametrization
        if self.curIter_ % self.lowerBoundComputationInterval == 0:
            ret.computeLowerBound = True
        return ret

    def end(self, lower_bound, upper_bound):  # :192-199
        ms = int((time.monotonic() - self.beginTime_) * 1000)
        self._say(f"final lower bound = {lower_bound:.10g}, upper bound = {upper_bound:.10g}")
        self._say(f"Optimization took {ms} milliseconds and {self.curIter_} iterations.")


class Solver:
    """Solver<LP_TYPE,VISITOR>, include/solver.hxx:29-372 (problem constructors / file IO out of scope)."""

    def __init__(self, lp, visitor: StandardVisitor | None = None):
        self.lp_ = lp
        self.visitor_ = visitor or StandardVisitor()
        self.lowerBound_ = -math.inf
        self.bestPrimalCost_ = math.inf
        self._solution = None      # best labeling on the host, or None while it only exists as the LP's kept (device-side) copy
        self._kept = False
        self.iter = 0
        self.history = []  # (lower bound, best primal, LpControl of that iteration) per visit

    def GetLP(self):
        return self.lp_

    @property
    def solution_(self):
        """best labeling registered so far (Solver::solution_, include/solver.hxx:340); downloaded on first use"""
        if self._solution is None and self._kept:
            self._solution = self.lp_.kept_labeling()
        return self._solution

    def lower_bound(self):
        return self.lowerBound_

    def primal_cost(self):
        return self.bestPrimalCost_

    def Solve(self):  # :230-257
        self.Begin()
        c = self.visitor_.begin(self.lp_)
        while not c.end and not c.error:
            self.PreIterate(c)
            self.Iterate(c)
            self.PostIterate(c)
            self.history.append((self.lowerBound_, self.bestPrimalCost_, c))
            c = self.visitor_.visit(c, self.lowerBound_, self.bestPrimalCost_)
            self.iter += 1
        if not c.error:
            self.End()
            self.RegisterPrimal()
            self.lowerBound_ = self.lp_.LowerBound()
            self.visitor_.end(self.lowerBound_, self.bestPrimalCost_)
        _ = self.solution_      # the solution outlives the LP handle
        return not c.error

    def Begin(self):  # :261-264
        self.lp_.Begin()

    def PreIterate(self, c):  # :267-270
        self.lp_.set_reparametrization(c.repam)

    def Iterate(self, c):  # :273-275
        self.lp_.ComputePass(self.iter)

    def PostIterate(self, c):  # :278-287
        if c.computeLowerBound:
            self.lowerBound_ = self.lp_.LowerBound()
            assert math.isfinite(self.lowerBound_)

    def End(self):  # :299-308
        self.lp_.End()

    def RegisterPrimal(self):  # :321-340
        cost = self.lp_.EvaluatePrimal()
        if cost < self.bestPrimalCost_ and self.lp_.CheckPrimalConsistency():
            self.bestPrimalCost_ = cost
            if hasattr(self.lp_, "keep_labeling"):      # device-side copy: the iteration does not wait for N labels to travel
                self.lp_.keep_labeling()
                self._solution, self._kept = None, True
            else:
                self._solution = self.lp_.labeling()


class MpRoundingSolver(Solver):
    """MpRoundingSolver<SOLVER>, include/solver.hxx:381-400: rounding interleaved with message passing."""

    def Iterate(self, c):
        if c.computePrimal:
            self.lp_.ComputeForwardPassAndPrimal(self.iter)
            self.RegisterPrimal()
            self.lp_.ComputeBackwardPassAndPrimal(self.iter)
            self.RegisterPrimal()
        else:
            super().Iterate(c)
