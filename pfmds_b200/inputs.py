"""Seeded synthetic inputs for the configurations of BASELINE.json and writers for the
reference's file formats: extended-xyz (md_read_write.f90:42-61), the fixed-order settings file
(md_simulation.f90:48-93) and the per-potential parameter files (INTERACTION_POTENTIALS/*.f90 read_*).

A *case* is a plain dict, consumed by `pfmds_b200.engine.configure` (C-ABI calls) and by `write_case`
(files for the run_md_simulation hosts):
  box(3) pos(N,3) vel(N,3) mass(N) names[N]  groups[[type names per group row]]  roles{...}
  integrators[(name,dt,len,snap,log)] ms_de nhc[(group,T,M,q1)] zero_momentum_period invert_z_vel
  interactions[{name, params[...], lists[(g1,g2,max,r_cut,period)]}]
  changes[(group_from, group_to, ts1, ts2, frec)]   optional: deposition entries (md_simulation.f90:63-71)
"""
from __future__ import annotations

import os

import numpy as np

COEF = 1.3806488 / 1.6605389217 * 1.0e-6  # (A/fs)^2 amu / K, md_general.f90:144
MASS_COEF = 1.6605389217 / 1.6021765654 * 100.0
KB = 1.3806488 / 1.6021765654 * 1.0e-4

# literature / builder-chosen parameter sets
RJL_CU = [0.0855, 1.224, 10.960, 2.278, 2.556, 5.5, 6.0]           # Cleri-Rosato Cu: A0 xi p q r0 / R1 R2
TB_BRENNER_I = [6.325, 1.29, 1.5, 1.315, 0.80469, 0.011304, 19.0, 2.5, 1.7, 2.0]  # d s b r0 delt a0 c0 d0 / R1 R2
LJC_C_CU = [0.02, 3.0, 2.0, 6.0, 7.0, 0.0]                         # eps sig delt / R1 R2 / simplified
MORSEC_C_CU = [0.03, 3.2, 1.2, 2.0, 6.0, 7.0, 0.0]                 # d r a delt / R1 R2 / simplified
# rebosc: A Q alpha / B(3) / beta(3) / T / g(6) / R1 R2 (REBOsolidcarbon.f90:12-25).  Pair terms: Brenner 2002 C-C; g: builder-chosen
# quintic through the knots of Brenner's G_C(cos) (-1, -2/3, -1/2, -1/3, 1) and G(0)=0.3
REBOSC_C = [10953.544162170, 0.3134602960833, 4.7465390606595, 12388.79197798, 17.56740646509, 30.71493208065,
            4.7204523127, 1.4332132499, 1.3826912506, -0.004048375,
            0.3, 1.00559171667, 1.751674625, 2.24048133333, 1.943325375, 0.75892695, 1.7, 2.0]


def maxwell(rng, mass, temperature, moving=None):
    """Maxwell velocities in A/fs at `temperature`, centre-of-mass motion removed, rescaled to the exact T."""
    n = len(mass)
    v = rng.standard_normal((n, 3)) * np.sqrt(COEF * temperature / mass)[:, None]
    sel = np.ones(n, bool) if moving is None else np.asarray(moving, bool)
    v[~sel] = 0.0
    if temperature <= 0 or sel.sum() == 0:
        return np.zeros((n, 3))
    m = mass[sel]
    v[sel] -= (m[:, None] * v[sel]).sum(0) / m.sum()
    ke = (m * (v[sel] ** 2).sum(1)).sum() / 2 * MASS_COEF
    t = 2 * ke / KB / (3 * sel.sum())
    v[sel] *= np.sqrt(temperature / t)
    return v


def ab_gas(n_side=22, spacing=4.0, seed=12345, temperature=100.0, frac_b=0.125, steps=(10000, 10000, 10000), period=20,
           period_log=1000, cap_aa=64, cap_ab=24, cap_ba=64, cap_bb=24):
    """C1: two-component A/B Lennard-Jones gas of the README (lj A-B + lj1g A-A + lj1g B-B, nvt->nve->nvms)."""
    rng = np.random.default_rng(seed)
    n = n_side ** 3
    g = np.arange(n_side) * spacing
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + spacing / 2
    pos = pos + rng.uniform(-0.3, 0.3, pos.shape)
    nb = int(round(n * frac_b))
    perm = rng.permutation(n)
    is_b = np.zeros(n, bool)
    is_b[perm[:nb]] = True
    order = np.concatenate([np.where(~is_b)[0], np.where(is_b)[0]])  # file order A..., B... keeps groups index-monotone (Q3)
    pos = pos[order]
    names = ["A"] * (n - nb) + ["B"] * nb
    mass = np.array([1.0] * (n - nb) + [10.0] * nb)
    vel = maxwell(rng, mass, temperature)
    box = np.array([n_side * spacing] * 3)
    return dict(
        title="ab_gas", box=box, pos=pos, vel=vel, mass=mass, names=names,
        groups=[["A", "B"], ["A", "#"], ["B", "#"], ["#", "#"]],
        roles=dict(all_moving=1, xyz_moving=1, z_moving=4, all_atoms=1, traj_group=3, period_traj=10 ** 9),
        integrators=[("nvt", 0.5, steps[0], 500000, period_log), ("nve", 0.5, steps[1], 500000, period_log), ("nvms", 0.5, steps[2], 500000, period_log)],
        ms_de=1e-8, nhc=[(1, temperature, 3, 10000.0)], zero_momentum_period=1000000, invert_z_vel=False, initial_temperature=temperature,
        interactions=[
            dict(name="lj", file="parameters_LJ_A-B.txt", params=[0.014, 3.2, 6.0, 7.0], lists=[(2, 3, cap_ab, 7.5, period), (3, 2, cap_ba, 7.5, period)]),
            dict(name="lj1g", file="parameters_LJ_A-A.txt", params=[0.0103, 3.405, 6.0, 7.0], lists=[(2, 2, cap_aa, 7.5, period)]),
            dict(name="lj1g", file="parameters_LJ_B-B.txt", params=[0.02, 3.0, 6.0, 7.0], lists=[(3, 3, cap_bb, 7.5, period)]),
        ],
    )


def lj_fluid(n_side=100, spacing=4.0, seed=7, temperature=100.0, steps=1000, period=20, cap=64):
    """Single-species lj1g system on the C1 lattice/density, for the LJ atom-steps/s target at large N."""
    rng = np.random.default_rng(seed)
    n = n_side ** 3
    g = np.arange(n_side) * spacing
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3) + spacing / 2
    pos = pos + rng.uniform(-0.3, 0.3, pos.shape)
    mass = np.full(n, 1.0)
    vel = maxwell(rng, mass, temperature)
    return dict(
        title="lj_fluid", box=np.array([n_side * spacing] * 3), pos=pos, vel=vel, mass=mass, names=["A"] * n,
        groups=[["A"], ["#"]], roles=dict(all_moving=1, xyz_moving=1, z_moving=2, all_atoms=1, traj_group=2, period_traj=10 ** 9),
        integrators=[("nve", 0.5, steps, 10 ** 9, 10 ** 9)], ms_de=1e-8, nhc=[], zero_momentum_period=10 ** 9, invert_z_vel=False,
        initial_temperature=temperature,
        interactions=[dict(name="lj1g", file="parameters_LJ_A-A.txt", params=[0.0103, 3.405, 6.0, 7.0], lists=[(1, 1, cap, 7.5, period)])],
    )


def lj_deposition(n_side=6, n_layers=3, n_deposit=8, spacing=3.8, seed=11, temperature=80.0, steps=60, period=5, ts1=3, ts2=40, frec=4,
                  thermostat=True):
    """Row (f) of SURVEY.md 8: atom deposition through `change_group_num`.  A Lennard-Jones substrate (type S, n_side^2*n_layers
    atoms) with n_deposit atoms of type D parked above it; the moving / interacting / thermostatted group 3 = [S, D] starts
    with the size of the substrate group 2 and takes one D atom at step ts1 and then every frec steps (md_general.f90:82-94).
    all_atoms is the growing group too, so zero_forces, the momentum removal and the writers follow group%N."""
    rng = np.random.default_rng(seed)
    g = np.arange(n_side) * spacing
    lay = np.stack(np.meshgrid(g, g, np.arange(n_layers) * spacing, indexing="ij"), -1).reshape(-1, 3) + np.array([1.0, 1.0, 4.0])
    lay = lay + rng.uniform(-0.1, 0.1, lay.shape)
    box = np.array([n_side * spacing, n_side * spacing, 60.0])
    top = 4.0 + (n_layers - 1) * spacing
    k = np.arange(n_deposit)
    dep = np.stack([(1.7 + 2.9 * k) % box[0], (2.3 + 5.3 * k) % box[1], top + 4.5 + 0.9 * (k % 30)], -1)
    pos = np.concatenate([lay, dep])
    ns = len(lay)
    names = ["S"] * ns + ["D"] * n_deposit
    mass = np.array([39.948] * ns + [39.948] * n_deposit)
    vel = maxwell(rng, mass, temperature, np.array([True] * ns + [False] * n_deposit))
    vel[ns:] = np.array([0.0, 0.0, -0.004])  # the parked atoms fly towards the surface once they are released
    nhc = [(3, temperature, 3, 3 * ns * KB * temperature * 100.0 ** 2)] if thermostat else []
    return dict(
        title="lj_deposition", box=box, pos=pos, vel=vel, mass=mass, names=names,
        groups=[["S", "D"], ["S", "#"], ["S", "D"], ["#", "#"]],
        roles=dict(all_moving=3, xyz_moving=3, z_moving=4, all_atoms=3, traj_group=3, period_traj=20),
        integrators=[("nvt" if thermostat else "nve", 1.0, steps, 20, 10)], ms_de=1e-8, nhc=nhc, zero_momentum_period=7, invert_z_vel=False,
        initial_temperature=temperature, changes=[(2, 3, ts1, ts2, frec)],
        interactions=[dict(name="lj1g", file="parameters_LJ_Ar.txt", params=[0.0103, 3.405, 6.0, 7.0], lists=[(3, 3, 80, 7.5, period)])],
    )


def cu_fcc(ncell=63, a=3.615, seed=2, temperature=300.0, steps=1000, dt=2.0, period=20, jitter=0.0, cells=None, q1=None):
    """C2: Cu fcc crystal, rjl (Cleri-Rosato), NVT 300 K.  ncell=63 -> 1 000 188 atoms; ncell=10 -> 4000 (parity size)."""
    rng = np.random.default_rng(seed)
    cx, cy, cz = cells if cells is not None else (ncell, ncell, ncell)
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) + 0.25
    ii, jj, kk = np.meshgrid(np.arange(cx), np.arange(cy), np.arange(cz), indexing="ij")
    cell = np.stack([ii, jj, kk], -1).reshape(-1, 1, 3)
    pos = ((cell + basis[None]) * a).reshape(-1, 3)
    if jitter > 0:
        pos = pos + rng.uniform(-jitter, jitter, pos.shape)
    n = len(pos)
    mass = np.full(n, 63.546)
    vel = maxwell(rng, mass, temperature)
    if q1 is None:
        q1 = 3 * n * KB * temperature * 100.0 ** 2  # thermostat period ~100 fs
    return dict(
        title="cu_fcc", box=np.array([cx * a, cy * a, cz * a]), pos=pos, vel=vel, mass=mass, names=["CU"] * n,
        groups=[["CU"], ["#"]], roles=dict(all_moving=1, xyz_moving=1, z_moving=2, all_atoms=1, traj_group=2, period_traj=10 ** 9),
        integrators=[("nvt", dt, steps, 10 ** 9, 10 ** 9)], ms_de=1e-8, nhc=[(1, temperature, 3, q1)], zero_momentum_period=10 ** 9,
        invert_z_vel=False, initial_temperature=temperature,
        interactions=[dict(name="rjl", file="parameters_RJL_Cu.txt", params=list(RJL_CU), lists=[(1, 1, 100, 6.5, period)])],
    )


def graphene_on_cu(gr_cells=(27, 27), cu_cells=(26, 26), n_layers=6, n_fixed=2, lz=60.0, seed=3, temperature=300.0, interface="ljc",
                   steps=(1000, 1000), dt=1.0, period=10, jitter=0.02, simplified=False, rep=(1, 1), carbon="tb"):
    """C3: graphene on Cu(111), rectangular moire cell: tb (C) + ljc|morsec (C-Cu) + rjl (Cu), nvt then nvms.
    File order C, CU, CU_fixed keeps every multi-type group index-monotone."""
    rng = np.random.default_rng(seed)
    a = 3.615
    ann = a / np.sqrt(2.0)
    cu_x, cu_y = ann, ann * np.sqrt(3.0)
    ncx, ncy = cu_cells[0] * rep[0], cu_cells[1] * rep[1]
    ngx, ngy = gr_cells[0] * rep[0], gr_cells[1] * rep[1]
    box = np.array([ncx * cu_x, ncy * cu_y, lz])
    dz = a / np.sqrt(3.0)
    z0 = 5.0
    cu, cu_fixed = [], []
    ii, jj = np.meshgrid(np.arange(ncx), np.arange(ncy), indexing="ij")
    base = np.stack([ii.ravel() * cu_x, jj.ravel() * cu_y], -1)
    for k in range(n_layers):
        shift = np.array([0.0, (k % 3) * cu_y / 3.0])
        for b in (np.array([0.0, 0.0]), np.array([0.5 * cu_x, 0.5 * cu_y])):
            xy = base + b + shift + np.array([0.1, 0.1])
            xy[:, 0] %= box[0]
            xy[:, 1] %= box[1]
            layer = np.concatenate([xy, np.full((len(xy), 1), z0 + k * dz)], 1)
            (cu_fixed if k < n_fixed else cu).append(layer)
    gx, gy = box[0] / ngx, box[1] / ngy
    ii, jj = np.meshgrid(np.arange(ngx), np.arange(ngy), indexing="ij")
    gbase = np.stack([ii.ravel() * gx, jj.ravel() * gy], -1)
    zc = z0 + (n_layers - 1) * dz + 3.2
    cs = []
    for fx, fy in ((0.0, 0.0), (0.0, 1.0 / 3.0), (0.5, 0.5), (0.5, 5.0 / 6.0)):
        xy = gbase + np.array([fx * gx, fy * gy]) + np.array([0.3, 0.2])
        xy[:, 0] %= box[0]
        xy[:, 1] %= box[1]
        cs.append(np.concatenate([xy, np.full((len(xy), 1), zc)], 1))
    c = np.concatenate(cs)
    cu = np.concatenate(cu)
    cu_fixed = np.concatenate(cu_fixed) if cu_fixed else np.zeros((0, 3))
    pos = np.concatenate([c, cu, cu_fixed])
    if jitter > 0:
        pos[: len(c) + len(cu)] += rng.uniform(-jitter, jitter, (len(c) + len(cu), 3))
    pos[:, 0] %= box[0]
    pos[:, 1] %= box[1]
    names = ["C"] * len(c) + ["CU"] * len(cu) + ["CU_fixed"] * len(cu_fixed)
    mass = np.array([12.011] * len(c) + [63.546] * (len(cu) + len(cu_fixed)))
    moving = np.array([True] * (len(c) + len(cu)) + [False] * len(cu_fixed))
    vel = maxwell(rng, mass, temperature, moving)
    n_move = int(moving.sum())
    inter = dict(name="ljc", file="parameters_LJC_C-Cu.txt", params=list(LJC_C_CU)) if interface == "ljc" else dict(
        name="morsec", file="parameters_MorseC_C-Cu.txt", params=list(MORSEC_C_CU))
    inter["params"][-1] = 1.0 if simplified else 0.0
    inter["lists"] = [(1, 2, 64, 7.5, period), (2, 1, 96, 7.5, period), (1, 1, 3, 1.9, period)]
    return dict(
        title="graphene_on_cu", box=box, pos=pos, vel=vel, mass=mass, names=names,
        # 1 gC, 2 gCu, 3 gMove, 4 gAll, 5 empty
        groups=[["C", "#", "#"], ["CU", "CU_fixed", "#"], ["C", "CU", "#"], ["C", "CU", "CU_fixed"], ["#", "#", "#"]],
        roles=dict(all_moving=3, xyz_moving=3, z_moving=5, all_atoms=4, traj_group=5, period_traj=10 ** 9),
        integrators=[("nvt", dt, steps[0], 10 ** 9, 100), ("nvms", dt, steps[1], 10 ** 9, 100)], ms_de=1e-8,
        nhc=[(3, temperature, 3, 3 * n_move * KB * temperature * 100.0 ** 2)], zero_momentum_period=10 ** 9, invert_z_vel=False,
        initial_temperature=temperature,
        interactions=[
            dict(name="tb", file="parameters_TB_C.txt", params=list(TB_BRENNER_I), lists=[(1, 1, 12, 2.6, period)]) if carbon == "tb" else
            dict(name="rebosc", file="parameters_REBOsc_C.txt", params=list(REBOSC_C), lists=[(1, 1, 12, 2.6, period)]),
            inter,
            dict(name="rjl", file="parameters_RJL_Cu.txt", params=list(RJL_CU), lists=[(2, 2, 100, 6.5, period)]),
        ],
    )


def graphene_rebosc(cells=(4, 3), lz=20.0, seed=5, temperature=300.0, steps=50, dt=0.5, period=5, jitter=0.04, with_tb=False):
    """Row (f) of SURVEY.md 8: a free-standing graphene sheet under `rebosc` (energy only in the reference; forces by central
    differences on truncated lists, md_interactions.f90:273-311).  Rectangular cell 2.46 x 4.2609 A, 4 C per cell."""
    rng = np.random.default_rng(seed)
    gx, gy = 2.46, 2.46 * np.sqrt(3.0)
    ii, jj = np.meshgrid(np.arange(cells[0]), np.arange(cells[1]), indexing="ij")
    base = np.stack([ii.ravel() * gx, jj.ravel() * gy], -1)
    cs = []
    for fx, fy in ((0.0, 0.0), (0.0, 1.0 / 3.0), (0.5, 0.5), (0.5, 5.0 / 6.0)):
        xy = base + np.array([fx * gx, fy * gy]) + np.array([0.3, 0.2])
        cs.append(np.concatenate([xy, np.full((len(xy), 1), lz / 2)], 1))
    pos = np.concatenate(cs) + rng.uniform(-jitter, jitter, (4 * len(base), 3))
    box = np.array([cells[0] * gx, cells[1] * gy, lz])
    pos[:, 0] %= box[0]
    pos[:, 1] %= box[1]
    n = len(pos)
    mass = np.full(n, 12.011)
    vel = maxwell(rng, mass, temperature)
    inter = [dict(name="rebosc", file="parameters_REBOsc_C.txt", params=list(REBOSC_C), lists=[(1, 1, 12, 2.6, period)])]
    if with_tb:
        inter.append(dict(name="tb", file="parameters_TB_C.txt", params=list(TB_BRENNER_I), lists=[(1, 1, 12, 2.6, period)]))
    return dict(
        title="graphene_rebosc", box=box, pos=pos, vel=vel, mass=mass, names=["C"] * n,
        groups=[["C"], ["#"]], roles=dict(all_moving=1, xyz_moving=1, z_moving=2, all_atoms=1, traj_group=2, period_traj=10 ** 9),
        integrators=[("nve", dt, steps, 10 ** 9, 10)], ms_de=1e-8, nhc=[], zero_momentum_period=10 ** 9, invert_z_vel=False,
        initial_temperature=temperature, interactions=inter,
    )


def graphene_on_cu_small(**kw):
    """Parity sub-case of C3: 6x4 graphene cells stretched onto 6x4 Cu cells, 96 C + 288 Cu = 384 atoms."""
    kw.setdefault("lz", 40.0)
    return graphene_on_cu(gr_cells=(6, 4), cu_cells=(6, 4), **kw)


def group_indexes(case, g):
    """1-based atom numbers of group g (1-based) in the reference's order: type column first, file order second
    (md_general.f90:70-77)."""
    names = np.asarray(case["names"])
    out = []
    for t in case["groups"][g - 1]:
        out.extend((np.where(names == t)[0] + 1).tolist())
    return np.asarray(out, dtype=np.int32)


# ---------------------------------------------------------------------------------------------
# file writers
def _param_lines(name, p):
    f = lambda xs: " ".join(repr(float(x)) for x in xs)
    if name in ("lj", "lj1g"):
        return [f(p[0:2]), f(p[2:4])]
    if name == "ljc":
        return [f(p[0:3]), f(p[3:5]), "T" if p[5] else "F"]
    if name == "morsec":
        return [f(p[0:4]), f(p[4:6]), "T" if p[6] else "F"]
    if name == "tb":
        return [f(p[0:8]), f(p[8:10])]
    if name == "rjl":
        return [f(p[0:5]), f(p[5:7])]
    if name == "rebosc":
        return [f(p[0:3]), f(p[3:6]), f(p[6:9]), f(p[9:10]), f(p[10:16]), f(p[16:18])]
    raise ValueError(name)


def write_xyz(path, case):
    """7f27.16 rows like write_particle_group (md_read_write.f90:65-83) so restarts are exact."""
    b = case["box"]
    with open(path, "w") as f:
        f.write("%12d\n" % len(case["mass"]))
        f.write('Lattice=" %.10f 0.0 0.0 0.0 %.10f 0.0 0.0 0.0 %.10f " Properties=pos:R:3:vel:R:3:mass:R:1:species:S:1\n' % (b[0], b[1], b[2]))
        pos, vel, mass, names = case["pos"], case["vel"], case["mass"], case["names"]
        for i in range(len(mass)):
            f.write("".join("%27.16f" % v for v in (*pos[i], *vel[i], mass[i])) + "    " + names[i] + "\n")


def write_case(directory, case, settings="md_run_settings.txt", xyz="init.xyz", log="md_run.log", md_step_limit=None, new_velocities=False):
    """Write xyz + settings + parameter files in the reference's grammar (code truth, not the stale README order)."""
    os.makedirs(directory, exist_ok=True)
    write_xyz(os.path.join(directory, xyz), case)
    r = case["roles"]
    ntypes = len(case["groups"][0])
    total = sum(i[2] for i in case["integrators"])
    L = []
    L.append("md_step_limit: %d" % (total if md_step_limit is None else md_step_limit))
    L.append("logfilename: %s" % log)
    L.append("init_xyz_filename: %s" % xyz)
    L.append("new_velocities: %s" % ("T" if new_velocities else "F"))
    L.append("zero_momentum_period: %d" % case["zero_momentum_period"])
    L.append("particle_types_num: %d" % ntypes)
    L.append("groups_num: %d" % len(case["groups"]))
    for k, g in enumerate(case["groups"]):
        L.append("\t%d\t%s" % (k + 1, "\t".join(g)))
    L.append("all_moving_atoms_group_num: %d" % r["all_moving"])
    L.append("xyz_moving_atoms_group_num: %d" % r["xyz_moving"])
    L.append("z_moving_atoms_group_num: %d" % r["z_moving"])
    L.append("all_atoms_group_num: %d" % r["all_atoms"])
    L.append("traj_group_num: %d" % r["traj_group"])
    L.append("period_traj: %d" % r["period_traj"])
    L.append("change_group_num: %d" % len(case.get("changes", [])))
    for fr, to, ts1, ts2, frec in case.get("changes", []):
        L.append("group_change_from_to: %d %d" % (fr, to))
        L.append("change_ts1_ts2_freq: %d %d %d" % (ts1, ts2, frec))
    L.append("invert_z_vel: %s" % ("T" if case["invert_z_vel"] else "F"))
    L.append("integrators_num: %d" % len(case["integrators"]))
    L.append(" name    dt         len     snap     log")
    for nm, dt, ln, snap, lg in case["integrators"]:
        L.append(" %s %.5f %d %d %d" % (nm, dt, ln, snap, lg))
    L.append("ms_de: %.6e" % case["ms_de"])
    L.append("nhc_num: %d" % len(case["nhc"]))
    for g, t, m, q in case["nhc"]:
        L.append("%d %r %d %r" % (g, float(t), m, float(q)))
    L.append("initial_temperature: %r" % float(case["initial_temperature"]))
    L.append("interactions_num: %d" % len(case["interactions"]))
    for it in case["interactions"]:
        L.append("%s %s" % (it["name"], it["file"]))
        for g1, g2, mx, rc, per in it["lists"]:
            L.append("%d %d %d %r %d" % (g1, g2, mx, float(rc), per))
        with open(os.path.join(directory, it["file"]), "w") as f:
            f.write("\n".join(_param_lines(it["name"], it["params"])) + "\n")
    with open(os.path.join(directory, settings), "w") as f:
        f.write("\n".join(L) + "\n")
    return os.path.join(directory, settings)


def cu_fcc_slab(rank, world, cells_per_rank=(63, 63, 63), a=3.615, seed=2, temperature=300.0):
    """This rank's share of the Cu fcc crystal of `world * cells_per_rank[0]` x cells_per_rank[1] x cells_per_rank[2] cells,
    generated locally (the 10^8-atom crystal of BASELINE.json configs[3] cannot be built whole on every rank).
    Atom numbering and positions are those `cu_fcc(cells=(world*cx, cy, cz), jitter=0)` gives for the same atoms; velocities
    are Maxwell draws per rank, and the global momentum / temperature rescale needs two sums over ranks: the caller reduces
    `sums` (sum m v (3), sum m, sum m v^2) over the ranks and calls `finish_velocities`.
    Returns (global_index_1based, pos, vel_raw, mass, box, sums)."""
    cx, cy, cz = cells_per_rank
    gx = world * cx
    basis = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]]) + 0.25
    ii, jj, kk = np.meshgrid(np.arange(rank * cx, (rank + 1) * cx), np.arange(cy), np.arange(cz), indexing="ij")
    cell = np.stack([ii, jj, kk], -1).reshape(-1, 1, 3)
    pos = ((cell + basis[None]) * a).reshape(-1, 3)
    # file order of cu_fcc(): cell-major (i, j, k) with the 4 basis atoms innermost
    cell_id = (cell[:, 0, 0] * cy + cell[:, 0, 1]) * cz + cell[:, 0, 2]
    gid = (cell_id[:, None] * 4 + np.arange(4)[None]).reshape(-1) + 1
    n = len(pos)
    mass = np.full(n, 63.546)
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[int(gid[0]), 0, 0, 0]))  # reproducible per rank, independent of the others
    vel = rng.standard_normal((n, 3)) * np.sqrt(COEF * temperature / mass)[:, None]
    sums = np.concatenate([(mass[:, None] * vel).sum(0), [mass.sum()], [(mass * (vel ** 2).sum(1)).sum()]])
    box = np.array([gx * a, cy * a, cz * a])
    return gid.astype(np.int64), pos, vel, mass, box, sums


def finish_velocities(vel, mass, sums_all_ranks, n_total, temperature):
    """Remove the global centre-of-mass velocity and rescale to exactly `temperature` (same steps as maxwell())."""
    p, m = sums_all_ranks[:3], sums_all_ranks[3]
    vcm = p / m
    v = vel - vcm
    ke = (sums_all_ranks[4] - m * (vcm ** 2).sum()) / 2 * MASS_COEF   # sum m (v - vcm)^2 = sum m v^2 - M vcm^2
    t = 2 * ke / KB / (3 * n_total)
    return v * np.sqrt(temperature / t)


def write_fitting_inputs(directory, case_a, case_b, interface="ljc", out_prefix="fit_", p_min=None, p_max=None, z=0.0, be0=-0.05, grd0=(0.05, 0.05),
                         zero_level=(0.0, 0.0)):
    """Inputs of run_gr_moire_fitting (fit_gr_moire.f90:107-182): two cells (settings + start xyz each), the bracket of the
    ljc / morsec parameters (min and max parameter files in parameter-file order) and the targets.  Returns the path of the
    fitting-parameters file (`-fpfn`)."""
    os.makedirs(directory, exist_ok=True)
    d = directory if directory.endswith(os.sep) else directory + os.sep
    names = []
    for tag, case in (("a", case_a), ("b", case_b)):
        write_case(d, case, settings="settings_%s.txt" % tag, xyz="cell_%s.xyz" % tag, log="md_%s.log" % tag)
        os.rename(d + "cell_%s.xyz" % tag, d + "start_%s.xyz" % tag)       # calc_error renames start -> xyz around every md()
        names.append(("settings_%s.txt" % tag, "start_%s.xyz" % tag, "final_%s.txt" % tag))
    it = [i for i in case_a["interactions"] if i["name"] == interface][0]
    p0 = list(it["params"])
    npar = 3 if interface == "ljc" else 4
    p_min = p_min if p_min is not None else [0.8 * v for v in p0[:npar]]
    p_max = p_max if p_max is not None else [1.2 * v for v in p0[:npar]]
    for nm, pv in (("params_min.txt", p_min), ("params_max.txt", p_max)):
        with open(d + nm, "w") as f:
            f.write(" ".join(repr(float(v)) for v in pv) + "\n")
            f.write("%r %r\n" % (float(p0[npar]), float(p0[npar + 1])))
    n_c = [sum(1 for t in c["names"] if t == "C") for c in (case_a, case_b)]
    L = []
    L.append("%-16s%s" % ("input_path:", d))
    L.append("%-16s%s" % ("out_path:", d))
    L.append("interaction_name: %s" % interface)
    L.append("output_prefix: %s" % out_prefix)
    for k in range(2):
        L.append("settings_%d: %s" % (k + 1, names[k][0]))
        L.append("start_xyz_%d: %s" % (k + 1, names[k][1]))
        L.append("c_num_%d: %d" % (k + 1, n_c[k]))
        L.append("zero_energy_%d: %r" % (k + 1, float(zero_level[k])))
        L.append("final_file_%d: %s" % (k + 1, names[k][2]))
    L.append("min_param_file: params_min.txt")
    L.append("max_param_file: params_max.txt")
    L.append("z: %r" % float(z))
    L.append("be0: %r" % float(be0))
    L.append("grd0_1: %r" % float(grd0[0]))
    L.append("grd0_2: %r" % float(grd0[1]))
    with open(d + "fitting.txt", "w") as f:
        f.write("\n".join(L) + "\n")
    return d + "fitting.txt"
