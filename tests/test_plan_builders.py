"""Index builders of the product library (vvb200_plan_create, O(N)) against the oracle's statement-by-statement
restatement of the reference's O(N^2) builders (VVIntegrator.cpp:123-155, CudaVVKernels.cpp:66-77, 483-594,
775-804, 884-891, 954-957, 1028-1031).  Integer arrays and fp64 DOFs / eta masses must be BIT-exact.
No GPU needed: plan creation makes no CUDA call."""
import dataclasses

import numpy as np
import pytest

INT_NAMES = ("particlesNH", "moleculesNH", "particleMolId", "drudePairs", "sortedByMol", "particlesInMolecules",
             "normalNH", "pairsNH", "normalLD", "pairsLD", "imagePairs", "electrolyte")
F64_NAMES = ("moleculeMasses", "moleculeInvMasses", "dof", "etaMass", "NkbT", "invMassTotal")


def compare(vv, vo, spec, params, precision="mixed"):
    plan = vv.Plan(spec, params, precision)
    oracle = vo.Oracle(spec, params, precision, literal=True)
    for name in INT_NAMES:
        a, b = plan.int_array(name), oracle.array(name)
        if name in ("normalLD", "pairsLD") and spec.langevin.size == 0:
            continue        # the reference never builds them without a Langevin kernel
        assert a.dtype == np.int32 and np.array_equal(a, b), name
    ng = oracle.num_temp_groups
    assert plan.num_temp_groups == ng
    for name in F64_NAMES:
        a, b = plan.f64_array(name), oracle.array(name)
        if name in ("etaMass", "NkbT"):
            b = b[: a.size]
        assert a.tobytes() == b.tobytes(), (name, a, b)
    return plan, oracle


@pytest.mark.parametrize("seed", range(8))
def test_ragged_topologies_bit_exact(vv, vo, seed):
    spec = vv.make_ragged(seed=seed)
    compare(vv, vo, spec, vv.Params().resolved_for(spec))


@pytest.mark.parametrize("com", [True, False])
@pytest.mark.parametrize("cmm", [True, False])
def test_bulk_dof_variants(vv, vo, com, cmm):
    spec = vv.make_bulk_ionic_liquid(30, hbond_constraints=True, has_cmm=cmm)
    params = dataclasses.replace(vv.Params(), use_com_temp_group=com)
    plan, _ = compare(vv, vo, spec, params)
    dof = plan.f64_array("dof")
    n_massive = int(np.sum(spec.masses > 0))
    # dof sum = 3 N_massive - N_constraints - 3 [CMMotionRemover]
    assert abs(dof.sum() - (3 * n_massive - spec.constraints.shape[0] - 3 * cmm)) < 1e-6


def test_nonpolar_single_group(vv, vo):
    spec = vv.make_nonpolar_box(64, 8)
    params = vv.Params().resolved_for(spec)
    plan, _ = compare(vv, vo, spec, params)
    assert plan.num_temp_groups == 1 and not params.use_com_temp_group and params.friction == 1.0


def test_edl_sets(vv, vo):
    spec = vv.make_edl(n_ion_pairs=10, n_electrode=90, electrode_molecules=3)
    plan, _ = compare(vv, vo, spec, vv.Params().resolved_for(spec))
    assert plan.random_request == 90 + 2           # padded request, CudaVVKernels.cpp:863 (SURVEY C-7)
    nh = plan.int_array("particlesNH")
    assert nh.min() == 90 and nh.size == 370       # neither Langevin nor image
    assert plan.tiled


def test_find_molecules_matches_oracle_and_scipy(vv, vo):
    rng = np.random.default_rng(5)
    n = 500
    bonds = rng.integers(0, n, size=(300, 2)).astype(np.int32)
    a, na = vv.find_molecules(n, bonds)
    b, nb = vo.find_molecules(n, bonds)
    from vvb200.system import molecules_from_bonds
    c, nc = molecules_from_bonds(n, bonds)
    assert na == nb == nc
    assert np.array_equal(a, b) and np.array_equal(a, c)
    # numbered by ascending first atom
    first = [int(np.flatnonzero(a == m)[0]) for m in range(na)]
    assert first == sorted(first)


def test_reference_exceptions_map_to_conflict(vv, vo):
    """VVIntegrator.cpp:149,155; CudaVVKernels.cpp:519,537,787,801 -> VVB200_ERR_CONFLICT with the reference's text"""
    base = vv.make_bulk_ionic_liquid(4)
    # Langevin on part of a molecule that is otherwise Nose-Hoover
    s1 = dataclasses.replace(base, langevin=np.array([0], np.int32)).finalize(mol_id=base.mol_id)
    with pytest.raises(vv.VVB200Error) as e:
        vv.Plan(s1, vv.Params())
    assert e.value.code == 2 and "same molecule" in e.value.message
    with pytest.raises(vo.OracleError):
        vo.Oracle(s1, vv.Params(), "mixed")
    # Langevin + cosine
    mol0 = np.flatnonzero(base.mol_id == 0).astype(np.int32)
    s2 = dataclasses.replace(base, langevin=mol0).finalize(mol_id=base.mol_id)
    with pytest.raises(vv.VVB200Error) as e:
        vv.Plan(s2, vv.Params(cos_acceleration=0.01))
    assert e.value.code == 2 and "periodic perturbation" in e.value.message
    # constraint straddling the two thermostats
    s3 = dataclasses.replace(s2, constraints=np.array([[0, 30]], np.int32)).finalize(mol_id=base.mol_id)
    with pytest.raises(vv.VVB200Error) as e:
        vv.Plan(s3, vv.Params())
    assert e.value.code == 2 and "Constrained particle pair" in e.value.message


def test_invalid_arguments(vv):
    spec = vv.make_bulk_ionic_liquid(2)
    with pytest.raises(vv.VVB200Error) as e:
        vv.Plan(spec, vv.Params(num_nh_chains=0))
    assert e.value.code == 1
    bad = dataclasses.replace(spec, drude_pairs=np.array([[0, 10 ** 6]], np.int32))
    with pytest.raises(vv.VVB200Error):
        vv.Plan(bad, vv.Params())


def test_tiles_cover_and_respect_units(vv):
    """fused-path tables: tiles partition [0,N), never split a thermostat molecule or a Drude pair, hold at
    most 512 slots and 128 molecules"""
    for spec in (vv.make_bulk_ionic_liquid(200), vv.make_edl(40, 300, 3), vv.make_nonpolar_box(300, 3),
                 vv.make_ragged(seed=3, scattered_molecules=0)):
        params = vv.Params().resolved_for(spec)
        plan = vv.Plan(spec, params)
        assert plan.tiled
        ts = plan.int_array("tileStart")
        assert ts[0] == 0 and ts[-1] == spec.n and np.all(np.diff(ts) > 0) and np.all(np.diff(ts) <= 512)
        tile_of = np.searchsorted(ts, np.arange(spec.n), side="right") - 1
        if spec.drude_pairs.size:
            assert np.array_equal(tile_of[spec.drude_pairs[:, 0]], tile_of[spec.drude_pairs[:, 1]])
        meta = plan.int_array("slotMeta")
        assert meta.dtype == np.uint32 and meta.size == spec.n
        local = meta & 0x7FF
        if params.use_com_temp_group:
            nh = plan.int_array("particlesNH")
            for m in np.unique(spec.mol_id[nh])[:200]:
                members = np.flatnonzero((spec.mol_id == m) & ((spec.masses > 0) | np.isin(np.arange(spec.n), nh)))
                assert len(set(tile_of[members])) == 1
            for t in range(len(ts) - 1):
                lm = local[ts[t]:ts[t + 1]]
                assert len(set(lm[lm != 0x7FF])) <= 128


def test_large_builder_is_linear_time(vv):
    """the reference's builders are O(N*M); ours must ingest a 1M-particle box in well under a second"""
    import time
    spec = vv.make_bulk_ionic_liquid(27648)
    t0 = time.perf_counter()
    plan = vv.Plan(spec, vv.Params().resolved_for(spec))
    dt = time.perf_counter() - t0
    assert plan.tiled and dt < 5.0
    assert plan.int_array("particlesNH").size == spec.n


@pytest.mark.parametrize("chains,chain_len,solvent", [(1, 330, 6), (3, 700, 40), (7, 1300, 500), (2, 5000, 0)])
def test_cut_molecules_fragment_tables(vv, chains, chain_len, solvent):
    """thermostat molecules longer than a tile: cut between Drude pairs only; every (tile, molecule) piece of a cut molecule
    is one fragment; a molecule's fragments are listed in tile order and together cover exactly its COM members; whole
    molecules stay inside one tile"""
    spec = vv.make_polymer(chains, chain_len, solvent, adjacent=True)
    params = vv.Params(max_drude_distance=0.02).resolved_for(spec)
    plan = vv.Plan(spec, params)
    assert plan.tiled
    ts = plan.int_array("tileStart")
    n_tiles = len(ts) - 1
    assert ts[0] == 0 and ts[-1] == spec.n and np.all(np.diff(ts) > 0) and np.all(np.diff(ts) <= 512)
    tile_of = np.searchsorted(ts, np.arange(spec.n), side="right") - 1
    assert np.array_equal(tile_of[spec.drude_pairs[:, 0]], tile_of[spec.drude_pairs[:, 1]])       # pairs are never cut
    off, mols, frag = plan.int_array("tileMolOffset"), plan.int_array("tileMolList"), plan.int_array("tileMolFrag")
    split, foff, flist = plan.int_array("splitMolId"), plan.int_array("splitFragOffset"), plan.int_array("splitFragList")
    assert off.size == n_tiles + 1 and off[-1] == mols.size == frag.size and np.all(np.diff(off) <= 128)
    assert split.size == chains and np.all(np.diff(split) > 0) and foff.size == split.size + 1 and foff[-1] == flist.size
    n_frag = int((frag >= 0).sum())
    assert sorted(flist.tolist()) == list(range(n_frag))                     # every fragment belongs to exactly one molecule
    meta = plan.int_array("slotMeta")
    local = (meta & 0x7FF).astype(np.int64)
    tile_of_entry = np.searchsorted(off, np.arange(mols.size), side="right") - 1
    frag_entry = {int(f): k for k, f in enumerate(frag) if f >= 0}
    for k, m in enumerate(split):
        fr = flist[foff[k]:foff[k + 1]]
        entries = [frag_entry[int(f)] for f in fr]
        assert all(mols[e] == m for e in entries)
        tiles = tile_of_entry[entries]
        assert np.all(np.diff(tiles) > 0) and len(tiles) >= 2                 # tile order, really cut
        members = np.flatnonzero(spec.mol_id == m)
        assert set(tile_of[members]) == set(tiles.tolist())                   # one fragment per tile the molecule touches
        # the slot words of the molecule's particles point at the molecule's entry of their tile
        for e in entries:
            t = tile_of_entry[e]
            inside = members[tile_of[members] == t]
            assert np.all(local[inside] == e - off[t])
    # whole (uncut) thermostat molecules: one tile, no fragment
    whole = np.setdiff1d(np.unique(mols), split)
    for m in whole[:300]:
        members = np.flatnonzero(spec.mol_id == m)
        assert len(set(tile_of[members])) == 1
    assert np.all(frag[np.isin(mols, whole)] == -1)


def test_image_of_table(vv):
    """fused image update: every parent's thread knows its image; absent when a parent has two images"""
    spec = vv.make_edl(40, 300, 3)
    params = vv.Params(mirror_location=1.0).resolved_for(spec)
    plan = vv.Plan(spec, params)
    image_of = plan.int_array("imageOf")
    assert image_of.size == spec.n
    assert np.array_equal(image_of[spec.image_pairs[:, 1]], spec.image_pairs[:, 0])
    assert (image_of >= 0).sum() == spec.image_pairs.shape[0]
    meta = plan.int_array("slotMeta")
    assert np.array_equal((meta >> 14) & 1, (image_of >= 0).astype(np.uint32))
    twice = dataclasses.replace(spec, image_pairs=np.concatenate([spec.image_pairs, spec.image_pairs[:1] + np.array([[1, 0]], np.int32)]))
    try:
        plan2 = vv.Plan(twice.finalize(mol_id=spec.mol_id) if hasattr(twice, "finalize") else twice, params)
        assert plan2.int_array("imageOf").size == 0                                            # falls back to the image kernel
    except vv.VVB200Error:
        pass    # the builder may refuse the doubled pair for other reasons (image inside a thermostat): also fine
