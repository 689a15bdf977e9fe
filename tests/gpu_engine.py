"""Adapter: drives the CUDA path (C ABI through quids_b200) with the packed states the checkers use."""
import numpy as np

import orc
import quids_b200 as qb

RULE_NAMES = {orc.RULE_HADAMARD: "hadamard", orc.RULE_ERASE_CREATE: "erase_create", orc.RULE_COIN: "coin", orc.RULE_SPLIT_MERGE: "split_merge"}
MOD_NAMES = {orc.MOD_CNOT: "cnot", orc.MOD_XGATE: "xgate", orc.MOD_YGATE: "ygate", orc.MOD_ZGATE: "zgate", orc.MOD_STEP: "step",
             orc.MOD_REVERSED_STEP: "reversed_step", orc.MOD_PHASE: "phase"}
NPARAMS = {orc.RULE_HADAMARD: 1, orc.RULE_ERASE_CREATE: 3, orc.RULE_COIN: 3, orc.RULE_SPLIT_MERGE: 3}
MOD_NPARAMS = {orc.MOD_CNOT: 2, orc.MOD_XGATE: 1, orc.MOD_YGATE: 1, orc.MOD_ZGATE: 1, orc.MOD_STEP: 0, orc.MOD_REVERSED_STEP: 0, orc.MOD_PHASE: 1}


class GpuEngine:
    """simulate/apply_modifier on packed states through the C ABI; `suffix` selects the registered
    variant of the rules ("" = fused symbolic hook, "_generic" = the four reference methods only)"""

    def __init__(self, suffix="", align=8):
        self.suffix = suffix
        self.align = align
        self.sym = qb.SymbolicIteration()
        self.last_phase_ms = None

    def _load(self, p: orc.Packed):
        it = qb.Iteration()
        it.upload_packed(p.sizes, p.mags, p.data, p.total_proba, align=self.align)
        return it

    @staticmethod
    def _store(it) -> orc.Packed:
        sizes, mags, data = it.download_packed()
        return orc.Packed(sizes, mags, data, it.total_proba)

    def rule(self, rid, params):
        return qb.Rule(RULE_NAMES[rid] + self.suffix, *list(params)[:NPARAMS[rid]])

    def simulate(self, p: orc.Packed, rid, params, k=orc.NO_TRUNCATION, tol=1e-30):
        qb.config.tolerance = tol
        qb.config.align_byte_length = self.align
        it, nxt = self._load(p), qb.Iteration()
        qb.simulate(it, self.rule(rid, params), nxt, self.sym, k)
        self.last_phase_ms = self.sym.phase_ms
        return self._store(nxt), self.sym.num_object, self.sym.num_object_after_interferences

    def apply_modifier(self, p: orc.Packed, mid, params=()):
        qb.config.align_byte_length = self.align
        it = self._load(p)
        qb.simulate(it, qb.Modifier(MOD_NAMES[mid], *list(params)[:MOD_NPARAMS[mid]]))
        return self._store(it)

    def hash_objects(self, p: orc.Packed, rid, params=(0, 0, 0)):
        qb.config.align_byte_length = self.align
        return self._load(p).hashes(self.rule(rid, params))
