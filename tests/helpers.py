"""Shared test helpers (CPU emulation of the kernel's work plan; tolerances)."""
import numpy as np

REL_TOL = 1e-6   # north-star parity bar: logp and gradient within 1e-6 relative (fp64 accumulation)


def grad_rel_err(g, g_ref):
    """max_i |g_i - ref_i| / max(|ref_i|, floor); floor = 1e-6 * max|ref| guards exact zeros."""
    g, g_ref = np.asarray(g), np.asarray(g_ref)
    floor = 1e-6 * max(np.max(np.abs(g_ref)), 1e-300)
    return float(np.max(np.abs(g - g_ref) / np.maximum(np.abs(g_ref), floor)))


def emulate_plan_gene_sums(plan, gid_sorted, values, G):
    """Per-gene sums of `values` computed the way lik_kernel + finalize_kernel assemble them:
    sequential accumulation per group; interior genes stored directly; partial genes through
    head/tail slots; groups of exclusive CTAs (one huge gene) reduced into one CTA slot; the split
    lists then sum the slots in order."""
    off, flags = plan["off"], plan["flags"]
    NG, gpc = len(flags), plan["groups_per_cta"]
    n_cta = NG // gpc
    slots = np.full(2 * NG + n_cta, np.nan)
    direct = np.full(G, np.nan)
    owners = np.zeros(G, dtype=np.int64)
    for v in range(NG):
        a, b = int(off[v]), int(off[v + 1])
        excl = bool(flags[v] & 4)
        cta = v // gpc
        if excl and np.isnan(slots[2 * NG + cta]):
            slots[2 * NG + cta] = 0.0
        if a >= b:
            continue
        if excl:
            assert gid_sorted[a] == gid_sorted[b - 1], "exclusive group spans genes"
            assert all(flags[cta * gpc:(cta + 1) * gpc] & 4), "exclusive flag must be CTA-uniform"
            slots[2 * NG + cta] += float(np.sum(values[a:b]))
            if flags[v] & 8:
                owners[gid_sorted[a]] += 1
            continue
        head, tail = bool(flags[v] & 1), bool(flags[v] & 2)
        g_first = gid_sorted[a]
        cur, acc = g_first, 0.0

        def flush(final, cur, acc):
            to_head = (cur == g_first) and head
            to_tail = (not to_head) and final and tail
            if to_head:
                slots[v] = acc
            elif to_tail:
                slots[NG + v] = acc
            else:
                assert np.isnan(direct[cur]), "gene written twice"
                direct[cur] = acc
            if not to_head:
                owners[cur] += 1
        for j in range(a, b):
            g = gid_sorted[j]
            if g != cur:
                flush(False, cur, acc)
                cur, acc = g, 0.0
            acc += values[j]
        flush(True, cur, acc)
    out = direct.copy()
    used = np.zeros(len(slots), dtype=bool)
    for k, g in enumerate(plan["split_gene"]):
        s = plan["split_slots"][plan["split_ptr"][k]:plan["split_ptr"][k + 1]]
        assert np.isnan(direct[g]), "split gene also written directly"
        assert not np.any(np.isnan(slots[s])), "split list reads an unwritten slot"
        assert not np.any(used[s]), "slot listed twice"
        used[s] = True
        out[g] = float(np.sum(slots[s]))
    assert np.array_equal(used, ~np.isnan(slots)), "a written slot is not read by any split list"
    return out, owners
