"""Synthetic input generators: the per-rank slab generator tiles to the same crystal as the global one."""
import numpy as np

from pfmds_b200 import inputs


def test_local_slab_generator_matches_the_global_crystal():
    world, cells = 3, (4, 3, 2)
    whole = inputs.cu_fcc(cells=(world * cells[0], cells[1], cells[2]), temperature=300.0)
    seen = np.zeros(len(whole["mass"]), bool)
    parts, sums = [], np.zeros(5)
    for r in range(world):
        gid, pos, vel, mass, box, s = inputs.cu_fcc_slab(r, world, cells)
        assert np.allclose(box, whole["box"])
        assert np.allclose(pos, whole["pos"][gid - 1], atol=1e-12)        # same numbering, same positions
        W = box[0] / world
        assert ((pos[:, 0] >= r * W) & (pos[:, 0] < (r + 1) * W)).all()    # and they are exactly this rank's slab
        assert not seen[gid - 1].any()
        seen[gid - 1] = True
        parts.append((vel, mass))
        sums += s
    assert seen.all()
    n = len(whole["mass"])
    v = np.concatenate([inputs.finish_velocities(vel, mass, sums, n, 300.0) for vel, mass in parts])
    m = np.concatenate([mass for _, mass in parts])
    assert np.abs((m[:, None] * v).sum(0)).max() < 1e-9                      # zero total momentum
    ke = (m * (v ** 2).sum(1)).sum() / 2 * inputs.MASS_COEF
    assert abs(2 * ke / inputs.KB / (3 * n) - 300.0) < 1e-9                  # exactly 300 K


def test_generators_are_seeded_and_monotone():
    a, b = inputs.graphene_on_cu_small(), inputs.graphene_on_cu_small()
    assert np.array_equal(a["pos"], b["pos"]) and np.array_equal(a["vel"], b["vel"])
    for g in range(1, len(a["groups"]) + 1):
        idx = inputs.group_indexes(a, g)
        assert np.all(np.diff(idx) > 0)          # file order C, CU, CU_fixed keeps every group index-monotone
    c = inputs.ab_gas(n_side=6)
    assert c["names"].count("B") == round(0.125 * 216)
