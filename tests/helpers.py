"""Test helpers: an oracle-backed evaluator for the synthetic-data generator (tests only)."""
import numpy as np


class OracleEval:
    def __init__(self, vb, O, model, toas):
        self.O = O
        self.md = vb.ModelDescriptor(model, toas)

    def phases(self, x):
        hi, lo, f, _, _ = self.O.phases(self.md, x)
        return hi, lo, f

    def resids_and_Ninvdiag(self, x):
        return self.O.residuals(self.md, x)


def oracle_eval_factory(vb, O):
    return lambda model, toas: OracleEval(vb, O, model, toas)
