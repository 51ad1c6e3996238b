"""Parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot run 4096 x 4096 or
4096 replicas in seconds): round trips, determinism / exact resume, periodic-halo consistency, monotone energy at
T = 0, sharding independence, energy bookkeeping."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N = 4096


@pytest.fixture(scope="module")
def big(cuda):
    import torch
    from cylinder_b200 import ops
    par = ops.make_params(cuda, wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01)
    Zt = torch.tensor([N], dtype=torch.int32, device=cuda)
    psi = torch.zeros(1, N, N, dtype=torch.complex64, device=cuda)
    ops.lattice_init(psi, Zt, .1, 77, 0)
    return par, Zt, psi


def _energy(ops, ws, par, a, variant=0):
    import torch
    S = ws.row_moments(par)[None].contiguous()
    Zt = torch.tensor([ws.Z], dtype=torch.int32, device=S.device)
    return ops.field_energy(S, Zt, ws.T, par, [a], None, variant).item()


def test_c4_pack_unpack_round_trip_and_halo(cuda, big):
    """4096 x 4096: pack -> unpack is the identity; after sweeps every periodic halo copy equals its interior source."""
    import torch
    from cylinder_b200 import ops, _abi
    par, Zt, psi = big
    ws = ops.HbmWorkspace(N, N, torch.complex64, cuda)
    ws.pack(psi[0])
    assert torch.equal(ws.unpack(), psi[0])
    chain = ops.new_chain(1, cuda, amplitude=.3)[0]
    ws.sweep(par, chain, 5, 0, 3, 0, _abi.SWEEP_ADAPT_SIGMA)
    lay = ws.layout
    hr, hc, pitch = lay["halo_rows"], lay["halo_cols"], lay["pitch"]
    rows = N + 2 * hr
    buf = ws.work[ws.cur.value * lay["buffer_elems"]:(ws.cur.value + 1) * lay["buffer_elems"]].reshape(2, rows, pitch // 2)
    padded = torch.empty(rows, pitch, dtype=torch.complex64, device=cuda)
    padded[:, 0::2], padded[:, 1::2] = buf[0], buf[1]                     # column-parity planes -> padded lattice
    inner = padded[hr:hr + N, hc:hc + N]
    assert torch.equal(inner, ws.unpack())
    assert torch.equal(padded[:hr, hc:hc + N], inner[-hr:]) and torch.equal(padded[hr + N:, hc:hc + N], inner[:hr])
    assert torch.equal(padded[hr:hr + N, :hc], inner[:, -hc:]) and torch.equal(padded[hr:hr + N, hc + N:hc + N + hc], inner[:, :hc])
    assert torch.equal(padded[:hr, :hc], inner[-hr:, -hc:]) and torch.equal(padded[hr + N:, hc + N:hc + N + hc], inner[:hr, :hc])


def test_c4_deterministic_and_exactly_resumable(cuda, big):
    """Counter-based RNG: the same (seed, sweep range) gives bit-identical lattices; 2 + 3 sweeps == 5 sweeps."""
    import torch
    from cylinder_b200 import ops, _abi
    par, Zt, psi = big
    outs = []
    for split in ((5,), (2, 3), (5,)):
        ws = ops.HbmWorkspace(N, N, torch.complex64, cuda)
        ws.pack(psi[0])
        chain = ops.new_chain(1, cuda, amplitude=.3)[0]
        s0 = 0
        for n in split:
            ws.sweep(par, chain, 9, s0, n, 0, _abi.SWEEP_ADAPT_SIGMA)
            s0 += n
        outs.append((ws.unpack().clone(), chain.clone()))
        del ws
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[2][0])
    assert torch.equal(outs[0][1], outs[1][1])
    assert 0.3 < (outs[0][1][_abi.CH_SITE_ACCEPTED] / outs[0][1][_abi.CH_SITE_PROPOSED]).item() < 0.9
    assert not torch.equal(outs[0][0], psi[0])


def test_c4_energy_never_increases_at_zero_temperature(cuda, big):
    """T = 0 accepts only dE <= 0 (accept rule B.2).  With a = 0 the local dE equals the change of the global energy
    (SURVEY F5), and same-colour moves touch disjoint bonds, so the total field energy is non-increasing sweep by sweep."""
    import torch
    from cylinder_b200 import ops, _abi
    _, Zt, psi = big
    par0 = ops.make_params(cuda, wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=0.0)
    ws = ops.HbmWorkspace(N, N, torch.complex64, cuda)
    ws.pack(psi[0])
    chain = ops.new_chain(1, cuda, amplitude=0.0, sigma_site=.02)[0]
    e = [_energy(ops, ws, par0, 0.0)]
    for s in range(4):
        ws.sweep(par0, chain, 3, s, 1, 0, 0)
        e.append(_energy(ops, ws, par0, 0.0))
    assert all(e[i + 1] <= e[i] + 1e-7 * abs(e[i]) for i in range(4)), e
    assert e[-1] < e[0]


@pytest.mark.parametrize("n", [4096, 4095, 2080, 6144])
def test_c4_measured_energy_and_profiles_match_independent_sums(cuda, n):
    """The fp32 C4 path with CYL_SWEEP_MEASURE at full size: the field energy and the profile sums that the sweep kernel takes
    from its finished tiles (tiles away from the lattice's end) or inside its colour passes (tiles cut by the end; with three
    colours all edge tiles) equal an independent evaluation of the swept lattice -- row moments in fp64 + the energy weights, and
    torch row sums -- sweep by sweep.  4096: two colours, every tile complete; 4095: three colours, last tile row / column cut;
    2080: two colours with a cut last tile column; 6144: more than 8192 tiles, where eight CTAs of the decision kernel share the
    tiles' partial sums and the last one to file its share decides.  Tolerance: fp32 partial sums of 16 sites per thread and 128 per tile row
    (2^-24 relative each) under terms of mixed sign: 2e-5 of the energy, 1e-5 of sum |psi|, 2e-3 absolute on the row sums of psi."""
    import torch
    from cylinder_b200 import ops, _abi
    par = ops.make_params(cuda, wavenumber=1.0, radius=1.0, alpha=-1.0, u=1.0, C=1.0, n=6.0, temperature=.01)
    psi = torch.zeros(1, n, n, dtype=torch.complex64, device=cuda)
    ops.lattice_init(psi, torch.tensor([n], dtype=torch.int32, device=cuda), .1, 5, 0)
    ws = ops.HbmWorkspace(n, n, torch.complex64, cuda)
    ws.pack(psi[0])
    del psi
    a = 0.3
    chain = ops.new_chain(1, cuda, amplitude=a)[0]
    obs, prof = ops.new_obs(1, n, cuda)
    e_sum, p2_sum = 0.0, 0.0
    prof_ref = torch.zeros(n, 3, dtype=torch.float64, device=cuda)
    nsw = 3
    for s in range(nsw):
        ws.sweep(par, chain, 11, s, 1, 0, _abi.SWEEP_ADAPT_SIGMA | _abi.SWEEP_MEASURE, obs=obs[0], prof=prof[0])
        lat = ws.unpack().to(torch.complex128)
        e_sum += _energy(ops, ws, par, a)
        p2_sum += (lat.real ** 2 + lat.imag ** 2).mean().item()
        prof_ref += torch.stack([lat.real.sum(1), lat.imag.sum(1), lat.abs().sum(1)], 1)
        del lat
    o = obs[0].cpu().numpy()
    assert o[_abi.OB_COUNT] == nsw
    assert o[_abi.OB_E_FIELD] == pytest.approx(e_sum, rel=2e-5)
    assert o[_abi.OB_PSI2] == pytest.approx(p2_sum, rel=1e-5)
    assert chain[_abi.CH_E_FIELD].item() == pytest.approx(_energy(ops, ws, par, a), rel=2e-5)
    pr, ref = prof[0].cpu().numpy(), prof_ref.cpu().numpy()
    np.testing.assert_allclose(pr[:, 2], ref[:, 2], rtol=1e-5)
    np.testing.assert_allclose(pr[:, :2], ref[:, :2], rtol=0, atol=2e-3)


def test_c3_batch_is_independent_of_batching(cuda):
    """4096 replicas (C3's grid): replica r gives the same chain whether it runs inside the full batch or inside a
    shard that starts at r0 (replica0 = r0) -- the property multi-GPU sharding relies on; identical parameters with
    different replica indices give different chains."""
    import sys, os
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from cylinder_b200.replicas import ReplicaBatch
    full = ReplicaBatch(bench.grid_params(4096, 0), dims=(50, 50), dtype=torch.complex64, device=cuda, seed=0x5EED)
    full.run(6)
    lo, hi = 1000, 1512
    shard = ReplicaBatch(bench.grid_params(hi - lo, lo), dims=(50, 50), dtype=torch.complex64, device=cuda, seed=0x5EED, replica0=lo)
    shard.run(6)
    zs = shard.Zmax
    assert torch.equal(shard.psi, full.psi[lo:hi, :zs]) and torch.equal(shard.chain, full.chain[lo:hi])
    assert torch.equal(shard.obs, full.obs[lo:hi])
    twin = ReplicaBatch(bench.grid_params(8192, 0), dims=(50, 50), dtype=torch.complex64, device=cuda, seed=0x5EED)
    twin.run(2)
    assert not torch.equal(twin.psi[:4096], twin.psi[4096:])            # same grid point, different replica index
    o = full.observables()
    assert torch.isfinite(o["field_energy"]).all() and (full.chain[:, 0].abs() < .8).all()
    assert int(full.Z_host.min()) == 40 and int(full.Z_host.max()) == 100 and (full.Z_host % 2 == 1).any()


def test_c3_longest_first_assignment_changes_nothing(cuda):
    """A launch of many replicas and >= 32 sweeps hands the replicas to the CTAs in descending row count (the kernel's own
    ranking of Z); CYL_SWEEP_INDEX_ORDER, a small shard or a short launch keep the index order.  Same results either
    way, replica by replica -- lattice, chain, observables and profiles each in the replica's own slot."""
    import sys, os
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from cylinder_b200.replicas import ReplicaBatch
    B = 2000                                                             # > 8 x SM count: ranked launch
    ranked = ReplicaBatch(bench.grid_params(B, 0), dims=(50, 50), dtype=torch.complex64, device=cuda, seed=11)
    ranked.run(40)
    plain = ReplicaBatch(bench.grid_params(B, 0), dims=(50, 50), dtype=torch.complex64, device=cuda, seed=11)
    plain.run(40, index_order=True)
    assert torch.equal(ranked.psi, plain.psi) and torch.equal(ranked.chain, plain.chain)
    assert torch.equal(ranked.obs, plain.obs) and torch.equal(ranked.prof, plain.prof)
    short = ReplicaBatch(bench.grid_params(B, 0), dims=(50, 50), dtype=torch.complex64, device=cuda, seed=11)
    short.run(20); short.run(20)                                         # short launches: index order; same chains
    assert torch.equal(ranked.psi, short.psi) and torch.equal(ranked.chain[:, :3], short.chain[:, :3])
    lo, hi = 700, 900                                                    # a shard below the threshold, long launch
    shard = ReplicaBatch(bench.grid_params(hi - lo, lo), dims=(50, 50), dtype=torch.complex64, device=cuda, seed=11, replica0=lo)
    shard.run(40)
    assert torch.equal(shard.psi, ranked.psi[lo:hi, :shard.Zmax]) and torch.equal(shard.chain, ranked.chain[lo:hi])
    assert torch.equal(shard.obs, ranked.obs[lo:hi])
    assert len(set(ranked.Z_host.tolist())) > 8                          # ragged: the ranking is not the identity


def test_c3_energy_bookkeeping(cuda):
    """The field energy recorded by the fused in-kernel accumulation equals an independent evaluation (row moments +
    weights) of the lattice the kernel leaves behind, for every replica of a C3-like batch (fp32 storage, <= 2e-5)."""
    import sys, os
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from cylinder_b200 import ops, _abi
    from cylinder_b200.replicas import ReplicaBatch
    b = ReplicaBatch(bench.grid_params(512, 0), dims=(50, 50), dtype=torch.complex64, device=cuda, seed=4, amplitude=.2)
    b.run(20, amp_moves=False, measure=True)            # amplitude fixed: the last measurement refers to the final lattice
    S = ops.row_moments(b.psi, b.Z, b.params)
    E = ops.field_energy(S, b.Z, b.T, b.params, b.chain[:, _abi.CH_AMPLITUDE].contiguous(), None, 0)
    np.testing.assert_allclose(b.chain[:, _abi.CH_E_FIELD].cpu().numpy(), E.cpu().numpy(), rtol=2e-5, atol=1e-4)


def test_replica_batch_checkpoint_resume_is_exact(cuda, tmp_path):
    """state_dict() / load_state_dict(): a run interrupted, saved with torch.save, and resumed in a NEW batch continues
    bit-identically (counter-based RNG: nothing but (seed, replica, sweep index) defines the stream)."""
    import sys, os
    import torch
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    from cylinder_b200.replicas import ReplicaBatch
    P = bench.grid_params(96, 17)
    a = ReplicaBatch(P, dims=(50, 50), dtype=torch.complex64, device=cuda, seed=11, replica0=17)
    a.run(30)
    torch.save(a.state_dict(), tmp_path / "ckpt.pt")
    a.run(25)
    b = ReplicaBatch(P, dims=(50, 50), dtype=torch.complex64, device=cuda, seed=999, replica0=0, initialize=False)
    b.load_state_dict(torch.load(tmp_path / "ckpt.pt", weights_only=False))
    b.run(25)
    assert torch.equal(a.psi, b.psi) and torch.equal(a.chain, b.chain) and torch.equal(a.obs, b.obs) and torch.equal(a.prof, b.prof)
    assert a.sweeps_done == b.sweeps_done == 55
