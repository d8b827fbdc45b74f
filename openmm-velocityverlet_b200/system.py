"""Synthetic systems in the shapes the reference plugin sees.

A SystemSpec is what VVIntegrator::initialize reads from OpenMM's System / ContextImpl /
DrudeForce plus the integrator's own particle lists; a HostState is OpenMM's posq /
posqCorrection / velm / force arrays in their device layouts, as numpy.
Everything is vectorised so the 16M-65M particle boxes of the scaling sweep build in seconds.
"""
from dataclasses import dataclass, field

import numpy as np

BOLTZ = 1.380649e-23 * 6.02214076e23 / 1000.0  # kJ/mol/K
E_FIELD_V_PER_NM = 1.60217662e-22 * 1.0  # kJ/(nm e) per V/nm (VVIntegrator.h:294)


def np_dtypes(precision):
    """(real, mixed) numpy dtypes of an OpenMM CudaPrecision mode."""
    return {"single": (np.float32, np.float32), "mixed": (np.float32, np.float64),
            "double": (np.float64, np.float64)}[precision]


def _i32(a, cols=None):
    a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
    return a.reshape(-1, cols) if cols else a.reshape(-1)


@dataclass
class SystemSpec:
    n: int
    masses: np.ndarray                      # [n] float64
    bonds: np.ndarray                       # [nb,2] int32 (all bonded pairs incl. Drude-parent, image-parent)
    drude_pairs: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))   # (drude, parent)
    constraints: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))
    has_cmm: bool = False
    langevin: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    image_pairs: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))   # (image, parent)
    electrolyte: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    mol_id: np.ndarray = None               # [n] int32; filled by finalize()
    n_mol: int = 0
    charges: np.ndarray = None              # [n] float64
    name: str = "system"

    @property
    def padded_n(self):
        return (self.n + 31) // 32 * 32     # OpenMM pads to a multiple of 32 [OMM-mem]

    def finalize(self, mol_id=None):
        self.masses = np.ascontiguousarray(self.masses, dtype=np.float64)
        self.bonds = _i32(self.bonds, 2)
        self.drude_pairs = _i32(self.drude_pairs, 2)
        self.constraints = _i32(self.constraints, 2)
        self.langevin = _i32(self.langevin)
        self.image_pairs = _i32(self.image_pairs, 2)
        self.electrolyte = _i32(self.electrolyte)
        if self.charges is None:
            self.charges = np.zeros(self.n)
        if mol_id is not None:
            self.mol_id = _i32(mol_id)
            self.n_mol = int(self.mol_id.max()) + 1
        else:
            # ContextImpl::getMolecules [OMM-mem]: every force's bonded pairs (DrudeForce: Drude-parent) AND the constraints
            links = np.concatenate([self.bonds, self.drude_pairs, self.constraints]) if self.n else self.bonds
            self.mol_id, self.n_mol = molecules_from_bonds(self.n, links)
        return self

    def c_arrays(self):
        return {"masses": self.masses, "mol_id": self.mol_id, "drude_pairs": self.drude_pairs,
                "constraints": self.constraints, "langevin": self.langevin, "image_pairs": self.image_pairs,
                "electrolyte": self.electrolyte}

    def subset_molecules(self, first_particle, last_particle):
        """Contiguous particle range [first,last) made of whole molecules -> a stand-alone SystemSpec
        (the per-rank partition of the multi-GPU sweep)."""
        a, b = int(first_particle), int(last_particle)
        sel = lambda pairs: pairs[(pairs[:, 0] >= a) & (pairs[:, 0] < b)] - a
        lst = lambda v: v[(v >= a) & (v < b)] - a
        mol = self.mol_id[a:b]
        s = SystemSpec(n=b - a, masses=self.masses[a:b].copy(), bonds=sel(self.bonds),
                       drude_pairs=sel(self.drude_pairs), constraints=sel(self.constraints), has_cmm=self.has_cmm,
                       langevin=lst(self.langevin), image_pairs=sel(self.image_pairs),
                       electrolyte=lst(self.electrolyte), charges=self.charges[a:b].copy(),
                       name=f"{self.name}[{a}:{b}]")
        return s.finalize(mol_id=mol - mol.min())


def molecules_from_bonds(n, bonds):
    """OpenMM's ContextImpl::getMolecules labelling [OMM-mem] (connected components numbered by
    ascending first atom), via scipy for speed."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    bonds = np.asarray(bonds).reshape(-1, 2)
    g = coo_matrix((np.ones(len(bonds), np.int8), (bonds[:, 0], bonds[:, 1])), shape=(n, n))
    _, lab = connected_components(g, directed=False)
    # relabel so that labels ascend with the first particle of each component
    first = np.full(lab.max() + 1, n, dtype=np.int64)
    np.minimum.at(first, lab, np.arange(n))
    order = np.argsort(first, kind="stable")
    rank = np.empty_like(order)
    rank[order] = np.arange(len(order))
    return rank[lab].astype(np.int32), int(lab.max() + 1)


@dataclass
class HostState:
    precision: str
    posq: np.ndarray            # [P,4] real
    corr: np.ndarray            # [P,4] real (mixed only) or None
    velm: np.ndarray            # [P,4] mixed
    force: np.ndarray           # [3,P] int64
    random: np.ndarray          # [R,4] float32
    box: tuple

    def copy(self):
        return HostState(self.precision, self.posq.copy(), None if self.corr is None else self.corr.copy(),
                         self.velm.copy(), self.force.copy(), self.random, self.box)

    def positions(self):
        """fp64 positions as the kernels reconstruct them: posq + posqCorrection."""
        x = self.posq[:, :3].astype(np.float64)
        if self.corr is not None:
            x = x + self.corr[:, :3].astype(np.float64)
        return x


# ---------------------------------------------------------------------------------------------
# topologies
# ---------------------------------------------------------------------------------------------
_CATION_PAIRS, _CATION_H, _ANION_PAIRS = 8, 11, 5
_ION_PAIR_SITES = 2 * _CATION_PAIRS + _CATION_H + 2 * _ANION_PAIRS   # 37


def _ion_pair_template(hbond_constraints):
    """One [C4mim-like cation + small anion] ion pair (statistics of examples/models/bulk_Im21):
    27-site cation = 8 adjacent (parent, Drude) pairs + 11 H, 10-site anion = 5 adjacent pairs,
    Drude right after its parent (oplspsffile.py:1515)."""
    masses, bonds, pairs, cons = [], [], [], []
    def add_pairs(base, count):
        for k in range(count):
            p, d = base + 2 * k, base + 2 * k + 1
            masses.extend([13.607 if k % 2 == 0 else 11.611, 0.4])
            pairs.append((d, p))
            bonds.append((p, d))
            if p != base:
                bonds.append((base, p))
    add_pairs(0, _CATION_PAIRS)
    for h in range(_CATION_H):
        i = 2 * _CATION_PAIRS + h
        masses.append(1.008)
        heavy = 2 * (h % _CATION_PAIRS)
        bonds.append((heavy, i))
        if hbond_constraints:
            cons.append((heavy, i))
    add_pairs(2 * _CATION_PAIRS + _CATION_H, _ANION_PAIRS)
    return (np.array(masses), np.array(bonds, np.int32), np.array(pairs, np.int32),
            np.array(cons, np.int32).reshape(-1, 2))


def _tile_template(n_units, unit_sites, arr, offset=0):
    """Replicate index array `arr` (values < unit_sites) for n_units consecutive units."""
    if arr.size == 0:
        return arr.reshape(0, arr.shape[1] if arr.ndim == 2 else 0)
    base = (np.arange(n_units, dtype=np.int64) * unit_sites + offset)
    out = arr[None, ...].astype(np.int64) + base.reshape((-1,) + (1,) * arr.ndim)
    return out.reshape((-1,) + arr.shape[1:]).astype(np.int32)


def make_bulk_ionic_liquid(n_ion_pairs, hbond_constraints=False, has_cmm=False, name=None):
    """Drude-polarizable ionic-liquid bulk (BASELINE configs 2/4/5): 37 particles, 13 Drude pairs and
    2 molecules per ion pair, all thermostatted by (TG)NH."""
    tm, tb, tp, tc = _ion_pair_template(hbond_constraints)
    n = n_ion_pairs * _ION_PAIR_SITES
    site = np.arange(_ION_PAIR_SITES)
    mol_in_unit = (site >= 2 * _CATION_PAIRS + _CATION_H).astype(np.int32)
    mol_id = (np.arange(n_ion_pairs, dtype=np.int64)[:, None] * 2 + mol_in_unit[None, :]).reshape(-1).astype(np.int32)
    spec = SystemSpec(n=n, masses=np.tile(tm, n_ion_pairs), bonds=_tile_template(n_ion_pairs, _ION_PAIR_SITES, tb),
                      drude_pairs=_tile_template(n_ion_pairs, _ION_PAIR_SITES, tp),
                      constraints=_tile_template(n_ion_pairs, _ION_PAIR_SITES, tc) if tc.size else np.zeros((0, 2), np.int32),
                      has_cmm=has_cmm, name=name or f"bulk_il_{n_ion_pairs}ip")
    return spec.finalize(mol_id=mol_id)


def make_nonpolar_box(n_molecules=512, atoms_per_molecule=8, has_cmm=True):
    """BASELINE config 1: ~4k-atom non-polarizable box, plain NH (no Drude => COM group off)."""
    n = n_molecules * atoms_per_molecule
    a = np.arange(1, atoms_per_molecule, dtype=np.int32)
    tb = np.stack([np.zeros_like(a), a], axis=1)
    masses = np.tile(np.array([12.011, 1.008, 1.008, 15.999, 14.007, 1.008, 12.011, 1.008])[:atoms_per_molecule]
                     if atoms_per_molecule <= 8 else np.full(atoms_per_molecule, 12.011), n_molecules)
    mol_id = np.repeat(np.arange(n_molecules, dtype=np.int32), atoms_per_molecule)
    spec = SystemSpec(n=n, masses=masses, bonds=_tile_template(n_molecules, atoms_per_molecule, tb),
                      has_cmm=has_cmm, name=f"nonpolar_{n}")
    return spec.finalize(mol_id=mol_id)


def make_polymer(n_chains=3, chain_len=700, n_solvent=40, has_cmm=False, adjacent=False):
    """Polarizable polymer chains under NH: each chain is ONE molecule of 2*chain_len particles (backbone atom +
    its Drude).  adjacent=False: parents listed first then all Drudes, so partners sit chain_len slots apart --
    neither the molecule nor the pairs fit a tile: the any-topology path.  adjacent=True: every Drude follows its
    parent (what OpenMM's Drude builders produce), with one hydrogen per second backbone atom -- the molecule is longer
    than a tile but can be cut between pairs: the fused path with cross-tile centres of mass.  Plus a few small
    solvent molecules."""
    masses, bonds, pairs = [], [], []
    base = 0
    for _ in range(n_chains):
        if adjacent:
            prev, i = -1, base
            for k in range(chain_len):
                masses.extend([13.607 if k % 2 == 0 else 11.611, 0.4])
                pairs.append((i + 1, i))
                bonds.append((i, i + 1))
                if prev >= 0:
                    bonds.append((prev, i))
                prev = i
                i += 2
                if k % 2 == 1:
                    masses.append(1.008)
                    bonds.append((prev, i))
                    i += 1
            base = i
            continue
        for k in range(chain_len):
            masses.append(13.607 if k % 2 == 0 else 11.611)
            if k:
                bonds.append((base + k - 1, base + k))
        for k in range(chain_len):
            masses.append(0.4)
            pairs.append((base + chain_len + k, base + k))
            bonds.append((base + k, base + chain_len + k))
        base += 2 * chain_len
    for _ in range(n_solvent):
        masses.extend([15.999, 1.008, 1.008])
        bonds.extend([(base, base + 1), (base, base + 2)])
        base += 3
    spec = SystemSpec(n=base, masses=np.array(masses), bonds=np.array(bonds, np.int32), drude_pairs=np.array(pairs, np.int32),
                      has_cmm=has_cmm, name=f"polymer_{n_chains}x{chain_len}{'_adjacent' if adjacent else ''}")
    return spec.finalize()


def make_edl(n_ion_pairs=511, n_electrode=2496, electrode_molecules=4, name=None, hbond_constraints=False):
    """BASELINE config 3 (examples/run-edl.py): Langevin electrode atoms, NH + external field on the
    electrolyte, one massless image particle per electrolyte particle bonded into its parent's
    molecule (run-edl.py:92-95) -- so electrolyte molecules are NOT contiguous in particle order."""
    tm, tb, tp, tc = _ion_pair_template(hbond_constraints)      # run-edl.py:30 constrains the H bonds of the ions
    n_el = n_ion_pairs * _ION_PAIR_SITES
    n = n_electrode + 2 * n_el
    per = n_electrode // electrode_molecules
    e_masses = np.where(np.arange(n_electrode) % 3 == 0, 95.937, 32.064)
    e_idx = np.arange(n_electrode, dtype=np.int32)
    e_bonds = np.stack([e_idx[:-1], e_idx[1:]], axis=1)
    e_bonds = e_bonds[(e_idx[1:] % per) != 0]                      # chains of `per` atoms
    ions = np.arange(n_el, dtype=np.int32) + n_electrode
    images = ions + n_el
    bonds = np.concatenate([e_bonds, _tile_template(n_ion_pairs, _ION_PAIR_SITES, tb, n_electrode),
                            np.stack([ions, images], axis=1)])
    masses = np.concatenate([e_masses, np.tile(tm, n_ion_pairs), np.zeros(n_el)])
    spec = SystemSpec(n=n, masses=masses, bonds=bonds,
                      drude_pairs=_tile_template(n_ion_pairs, _ION_PAIR_SITES, tp, n_electrode),
                      constraints=_tile_template(n_ion_pairs, _ION_PAIR_SITES, tc, n_electrode) if tc.size else np.zeros((0, 2), np.int32),
                      langevin=e_idx, image_pairs=np.stack([images, ions], axis=1), electrolyte=ions,
                      name=name or f"edl_{n_ion_pairs}ip")
    return spec.finalize()


def make_ragged(seed=0, n_molecules=60, max_size=40, drude_fraction=0.3, massless_fraction=0.05,
                langevin_molecules=3, images=5, electrolyte_dups=True, constraints=True, has_cmm=True,
                scattered_molecules=2, max_partner_gap=3):
    """Small adversarial topology for the edge cases: ragged molecule sizes (1..max_size), Drude
    partners that are not adjacent, massless sites, Langevin molecules, image particles bonded to
    parents (scattered molecules), duplicated electrolyte entries, constraints, CMMotionRemover."""
    rng = np.random.default_rng(seed)
    sizes = rng.integers(1, max_size + 1, size=n_molecules)
    sizes[rng.integers(0, n_molecules, size=max(1, n_molecules // 10))] = 1        # monatomic ions
    starts = np.concatenate([[0], np.cumsum(sizes)])
    n_real = int(starts[-1])
    masses = rng.choice([1.008, 12.011, 14.007, 15.999, 32.06], size=n_real)
    bonds, pairs, cons = [], [], []
    ld_mols = set(rng.choice(n_molecules, size=min(langevin_molecules, n_molecules), replace=False).tolist())
    in_pair = np.zeros(n_real, bool)
    for m in range(n_molecules):
        a, b = starts[m], starts[m + 1]
        for i in range(a + 1, b):
            bonds.append((int(rng.integers(a, i)), i))
        # Drude pairs: parent p, Drude d = p + gap (gap >= 1), both unused so far
        for p in range(a, b):
            if in_pair[p] or rng.random() > drude_fraction:
                continue
            d = p + int(rng.integers(1, max_partner_gap + 1))
            if d < b and not in_pair[d]:
                in_pair[p] = in_pair[d] = True
                masses[d] = 0.4
                masses[p] = max(masses[p] - 0.4, 0.6)
                pairs.append((d, p))
                bonds.append((p, d))
        if constraints:
            for i in range(a + 1, b):
                if masses[i] == 1.008 and not in_pair[i] and not in_pair[a] and rng.random() < 0.5:
                    cons.append((int(a), i))
    free = np.flatnonzero(~in_pair)
    free = np.setdiff1d(free, starts[:-1])      # keep one massive site per molecule (an all-massless NH
                                                # molecule is 0*inf = NaN in the reference's COM kernel)
    massless = rng.choice(free, size=int(massless_fraction * len(free)), replace=False) if len(free) else []
    masses[massless] = 0.0
    langevin = np.concatenate([np.arange(starts[m], starts[m + 1]) for m in sorted(ld_mols)]) if ld_mols else []
    langevin = np.array(langevin, dtype=np.int32)
    rng.shuffle(langevin)                       # addParticleLangevin order is the user's
    nh_particles = np.setdiff1d(np.arange(n_real), langevin)
    # images: massless copies of a few thermostatted particles, appended after all real particles
    parents = rng.choice(nh_particles, size=min(images, len(nh_particles)), replace=False) if images else np.array([], int)
    image_idx = n_real + np.arange(len(parents))
    for im, pa in zip(image_idx, parents):
        bonds.append((int(pa), int(im)))
    # a few more scattered molecules: bond a distant massless site into an earlier molecule
    extra = []
    for k in range(scattered_molecules):
        idx = n_real + len(parents) + k
        bonds.append((int(rng.choice(nh_particles)), idx))
        extra.append(idx)
    n = n_real + len(parents) + len(extra)
    masses = np.concatenate([masses, np.zeros(len(parents) + len(extra))])
    electrolyte = rng.choice(nh_particles, size=min(len(nh_particles), max(4, n_real // 3)), replace=False)
    if electrolyte_dups and len(electrolyte) > 2:
        electrolyte = np.concatenate([electrolyte, electrolyte[:2], electrolyte[:1]])
    # constraints must not straddle thermostats or touch massless/Drude particles
    cons = [c for c in cons if masses[c[0]] > 0 and masses[c[1]] > 0]
    spec = SystemSpec(n=n, masses=masses, bonds=np.array(bonds, np.int32).reshape(-1, 2),
                      drude_pairs=np.array(pairs, np.int32).reshape(-1, 2),
                      constraints=np.array(cons, np.int32).reshape(-1, 2), has_cmm=has_cmm, langevin=langevin,
                      image_pairs=np.stack([image_idx, parents], axis=1).astype(np.int32) if len(parents) else np.zeros((0, 2), np.int32),
                      electrolyte=electrolyte.astype(np.int32), name=f"ragged_{seed}")
    return spec.finalize()


# ---------------------------------------------------------------------------------------------
# state
# ---------------------------------------------------------------------------------------------
def make_state(spec, precision="mixed", temperature=333.0, drude_temperature=1.0, seed=12345,
               density=158.0, force_sigma=1000.0, n_random=0, drude_spread=0.005, mirror=None):
    """Positions uniform in a cubic box at `density` particles/nm^3, Drude = parent + N(0,drude_spread);
    Maxwell-Boltzmann velocities (Drude relative motion at drude_temperature); frozen forces
    N(0, force_sigma) kJ/mol/nm in OpenMM's 2^32 fixed point; charges; optional N(0,1) float4 stream."""
    real, mixed = np_dtypes(precision)
    n, P = spec.n, spec.padded_n
    n_box = max(n - spec.image_pairs.shape[0], 1)
    L = (n_box / density) ** (1.0 / 3.0)
    rng_x = np.random.Generator(np.random.PCG64(seed))
    rng_v = np.random.Generator(np.random.PCG64(seed + 1))
    rng_f = np.random.Generator(np.random.PCG64(seed + 2))
    rng_r = np.random.Generator(np.random.PCG64(seed + 3))
    rng_q = np.random.Generator(np.random.PCG64(seed + 4))

    x = rng_x.uniform(0.0, L, size=(n, 3))
    d, p = (spec.drude_pairs[:, 0], spec.drude_pairs[:, 1]) if spec.drude_pairs.size else (np.zeros(0, int), np.zeros(0, int))
    if d.size:
        x[d] = x[p] + rng_x.normal(0.0, drude_spread, size=(d.size, 3))
    if spec.image_pairs.size:
        im, pa = spec.image_pairs[:, 0], spec.image_pairs[:, 1]
        m = L / 2 if mirror is None else mirror
        x[im] = x[pa]
        x[im, 2] = 2 * m - x[pa, 2]

    masses = spec.masses
    massive = masses > 0
    v = np.zeros((n, 3))
    sig = np.zeros(n)
    sig[massive] = np.sqrt(BOLTZ * temperature / masses[massive])
    v = rng_v.normal(size=(n, 3)) * sig[:, None]
    if d.size:
        mu = masses[d] * masses[p] / (masses[d] + masses[p])
        v[d] = v[p] + rng_v.normal(size=(d.size, 3)) * np.sqrt(BOLTZ * drude_temperature / mu)[:, None]

    q = spec.charges.copy()
    if not np.any(q):
        q = rng_q.normal(0.0, 0.5, size=n)
        if d.size:
            q[d] = -(1.0 + np.abs(rng_q.normal(0.0, 0.3, size=d.size)))
        if spec.image_pairs.size:
            q[spec.image_pairs[:, 0]] = -q[spec.image_pairs[:, 1]]

    posq = np.zeros((P, 4), dtype=real)
    corr = None
    if precision == "mixed":
        hi = x.astype(np.float32)
        posq[:n, :3] = hi
        corr = np.zeros((P, 4), dtype=np.float32)
        corr[:n, :3] = (x - hi.astype(np.float64)).astype(np.float32)
    else:
        posq[:n, :3] = x.astype(real)
    posq[:n, 3] = q.astype(real)

    velm = np.zeros((P, 4), dtype=mixed)
    velm[:n, :3] = v.astype(mixed)
    inv = np.zeros(n)
    inv[massive] = 1.0 / masses[massive]
    velm[:n, 3] = inv.astype(mixed)          # OpenMM stores (mixed)(1/mass), 0 for massless

    f = rng_f.normal(0.0, force_sigma, size=(3, n))
    f[:, ~massive] = 0.0
    force = np.zeros((3, P), dtype=np.int64)
    force[:, :n] = (f * 4294967296.0).astype(np.int64)      # truncation toward zero, like (long long)

    random = rng_r.normal(size=(max(n_random, 1), 4)).astype(np.float32)
    return HostState(precision, posq, corr, velm, force, random, (L, L, L))
