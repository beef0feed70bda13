"""Shared-memory tile layout: every warp-wide access pattern the kernels use must be
bank-conflict free (rtl/ntt_coeff_banks.v's banking constraint, re-expressed for 32 banks x 4 B)."""
import pytest

import emu

#        word logn logr ppc   register-field positions used (forward and inverse passes)
GEOMS = [(4, 8, 4, 16, (4, 0)), (4, 10, 5, 8, (5, 0)), (4, 10, 4, 4, (6, 2, 0, 4)), (4, 12, 4, 1, (8, 4, 0)),
         (4, 12, 5, 2, (7, 2, 0, 5)), (4, 12, 3, 1, (9, 6, 3, 0)),
         (8, 12, 4, 1, (8, 4, 0)), (8, 12, 3, 1, (9, 6, 3, 0)), (8, 8, 4, 16, (4, 0)), (8, 10, 4, 4, (6, 2, 0, 4))]


def wavefronts(wb, slots):
    """Number of shared-memory wavefronts a warp request needs: 4-byte words are served 32 lanes
    at a time, 8-byte words 16 lanes at a time; a wavefront serves one address per bank."""
    group = 32 if wb == 4 else 16
    total = 0
    for g in range(0, 32, group):
        banks = {}
        for s in slots[g:g + group]:
            for word in range(wb // 4):
                banks.setdefault((s * (wb // 4) + word) % 32, set()).add(s)
        total += max(len(v) for v in banks.values())
    return total


@pytest.mark.parametrize("wb,logn,logr,ppc,fields", GEOMS)
def test_tile_accesses_are_conflict_free(wb, logn, logr, ppc, fields):
    L = emu.lib()
    p = 1 << (logn - logr)
    threads = p * ppc
    ideal = 1 if wb == 4 else 2
    for lo in fields:
        for warp in range(0, threads, 32):
            for k in range(1 << logr):
                slots = [L.emu_slot(wb, logn, logr, lo, t >> (logn - logr), t & (p - 1), k)
                         for t in range(warp, min(warp + 32, threads))]
                assert len(set(slots)) == len(slots)
                if len(slots) == 32:
                    assert wavefronts(wb, slots) == ideal, (lo, warp, k)


@pytest.mark.parametrize("wb,logn,logr,ppc,fields", GEOMS)
def test_tile_slot_map_is_a_bijection(wb, logn, logr, ppc, fields):
    L = emu.lib()
    p = 1 << (logn - logr)
    n = 1 << logn
    for lo in fields:
        seen = set()
        for pl in range(ppc):
            for t in range(p):
                for k in range(1 << logr):
                    seen.add(L.emu_slot(wb, logn, logr, lo, pl, t, k))
        assert seen == set(range(ppc * n))


#            word logn logr ppc  fields     patterns that are allowed one extra wavefront
PADDED = [(8, 12, 4, 1, (8, 4, 0), ()), (4, 8, 4, 16, (4, 0), ()), (4, 10, 5, 8, (5, 0), ()), (4, 12, 4, 1, (8, 4, 0), (8,))]


@pytest.mark.parametrize("wb,logn,logr,ppc,fields,lossy", PADDED)
def test_padded_tile_is_conflict_free_and_injective(wb, logn, logr, ppc, fields, lossy):
    """Cfg<..., PAD = 1>: slot = E + (E >> log2 R) (one pad word per R).  Every warp access of every layout is
    conflict-free, except the 32-bit N = 4096 shape's accesses with the register field at bit 8: lane 31 lands on lane
    0's bank (one extra wavefront), which the shape is measured to afford (kernels.cuh, Cfg)."""
    L = emu.lib()
    p = 1 << (logn - logr)
    threads = p * ppc
    ideal = 1 if wb == 4 else 2
    for lo in fields:
        seen = set()
        for warp in range(0, threads, 32):
            for k in range(1 << logr):
                slots = [L.emu_slot_ex(wb, logn, logr, lo, t >> (logn - logr), t & (p - 1), k, 1)
                         for t in range(warp, min(warp + 32, threads))]
                if len(slots) == 32:
                    assert wavefronts(wb, slots) == ideal + (1 if lo in lossy else 0), (lo, warp, k)
                seen.update(slots)
        n_all = ppc << logn
        assert len(seen) == n_all and max(seen) < n_all + (n_all >> logr)


@pytest.mark.parametrize("wb,logn,logr,ppc,fields", GEOMS)
def test_warp_local_exchanges_stay_inside_their_thread_group(wb, logn, logr, ppc, fields):
    """exchange_is_warp_local() in kernels.cuh: a regrouping between register fields at 0 and at LO <= 5 is ordered by
    __syncwarp() only.  That is sound iff every tile slot is written (layout A) and read (layout B) by threads of the
    same aligned group of 2^LO <= 32 consecutive threads -- for both directions and for every polynomial of the CTA."""
    L = emu.lib()
    p = 1 << (logn - logr)
    los = sorted(set(fields))
    for lo in los:
        if lo == 0 or lo > 5 or 0 not in los:
            continue
        group = 1 << lo
        owner = {}
        for layout in (0, lo):
            for pl in range(ppc):
                for t in range(p):
                    for k in range(1 << logr):
                        s = L.emu_slot(wb, logn, logr, layout, pl, t, k)
                        g = (pl * p + t) // group
                        assert owner.setdefault(s, g) == g, (lo, layout, pl, t, k)
        assert group <= 32 and 32 % group == 0


@pytest.mark.parametrize("logn,logr,cs", [(12, 3, 4), (14, 4, 4), (15, 4, 8), (12, 4, 2)])
def test_cluster_exchange_addressing(logn, logr, cs):
    """The distributed-shared-memory exchange of the cluster kernels (kernels.cuh cluster_exchange): every remote
    store stays inside the cluster and the destination buffer, no slot has two writers, and every reader finds the
    coefficient of its next layout -- for every regrouping a forward and an inverse transform perform."""
    L = emu.lib()
    npass = (logn + logr - 1) // logr
    fwd = [max(logn - (p + 1) * logr, 0) for p in range(npass)]
    inv = [min(p * logr, logn - logr) for p in range(npass)]
    for los in (fwd, inv):
        for lo_from, lo_to in zip(los, los[1:]):
            assert L.emu_cluster_exchange_violations(logn, logr, cs, lo_from, lo_to) == 0, (lo_from, lo_to)
    assert L.emu_cluster_exchange_violations(logn, logr, cs, fwd[0], fwd[0]) == 0      # identity regrouping
