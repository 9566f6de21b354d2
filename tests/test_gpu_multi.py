"""One process driving several GPUs (SURVEY.md section 8e, 8b "Threading"): ca_init(devices, ndev > 1) shards the label
matrix by 32-partition spans, every device packs its share and stores the compact labels into every device's memory, runs
the pairs of its own partitions, and one NCCL all-reduce of 64-bit integer sums combines the results.  The reference's
analogue is foreach %dopar% over pairs (R/ECS.R:953-959); the checker is the CPU oracle, and -- because the sums are
fixed-point integers -- the single-device result of the same library, BIT FOR BIT.

Needs at least two visible GPUs (gpurun --gpus 2); skipped otherwise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _ngpu():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


needs2 = pytest.mark.skipif(_ngpu() < 2, reason="needs two visible GPUs")


@pytest.fixture()
def two_devices():
    from clustassess_b200 import _capi

    _capi.lib().ca_shutdown()
    n = _capi.init(list(range(min(_ngpu(), 8))))
    yield n
    _capi.lib().ca_shutdown()
    _capi.init(None)


@needs2
def test_sharded_consistency_matches_oracle_and_single_device(two_devices):
    import oracle
    from clustassess_b200 import _capi, ecs, synth

    L, w = synth.config("C3", m=100, n=30_000)
    m, n = L.shape
    res = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    ecc_ref, sim_ref = oracle.weighted_element_consistency(L, w, calculate_sim_matrix=True)
    assert np.abs(res["ecc"] - ecc_ref).max() < TOL
    assert np.abs(res["ecs_matrix"] - sim_ref).max() < TOL
    # the same job on one device: identical bits (integer sums do not depend on who adds them)
    _capi.lib().ca_shutdown()
    _capi.init([0])
    one = ecs.weighted_element_consistency(list(L), weights=w, calculate_sim_matrix=True)
    assert np.array_equal(one["ecc"], res["ecc"])
    assert np.array_equal(one["ecs_matrix"], res["ecs_matrix"])


@needs2
@pytest.mark.parametrize("m", [33, 64, 65, 97, 130])
def test_sharded_span_boundaries(two_devices, m):
    """Partition counts around the 32-partition spans the devices share out (a partial last span, an odd number of spans)."""
    import oracle
    from clustassess_b200 import ecs

    rng = np.random.default_rng(500 + m)
    n = 5000
    base = rng.integers(0, 40, size=n)
    L = np.stack([np.where(rng.random(n) < 0.25, rng.integers(0, 40 + p % 9, size=n), base) for p in range(m)]).astype(np.int32)
    ecc = oracle.weighted_element_consistency(L)
    got = ecs.weighted_element_consistency(list(L))
    assert np.abs(got - ecc).max() < TOL
    sim = ecs.element_sim_matrix(list(L))
    assert np.abs(sim - oracle.element_sim_matrix(L)).max() < TOL


@needs2
def test_sharded_resident_set_and_matrix_form(two_devices):
    """A resident sharded set: upload once (matrix form and column form), run twice, subsets by index."""
    import oracle
    from clustassess_b200 import _capi, synth

    L, _ = synth.config("C4", m=96, n=8000)
    m, n = L.shape
    dl = _capi.DeviceLabels(n, m).upload(L)
    assert dl.sharded
    ecc1, sim1 = dl.consistency(want_matrix=True)
    ecc2, _ = dl.consistency()
    assert np.array_equal(ecc1, ecc2)  # reproducible bits
    ecc_ref, sim_ref = oracle.weighted_element_consistency(L, None, calculate_sim_matrix=True)
    assert np.abs(ecc1 - ecc_ref).max() < TOL and np.abs(sim1 - sim_ref).max() < TOL
    idx = [3, 70, 5, 41, 42, 43, 95, 0] + list(range(10, 60))
    sub, _ = dl.consistency(idx)
    assert np.abs(sub - oracle.weighted_element_consistency(L[idx])).max() < TOL
    dl.close()
    dc = _capi.DeviceLabels(n, m).upload_columns([np.ascontiguousarray(r) for r in L])
    ecc3, _ = dc.consistency()
    assert np.array_equal(ecc1, ecc3)
    dc.close()


@needs2
def test_element_consistency_dedup_route_on_several_devices(two_devices):
    """element_consistency = dedup on the primary device + consistency of the kept partitions on all devices."""
    import oracle
    from clustassess_b200 import ecs, synth

    L, _ = synth.config("C3", m=40, n=12_000)
    dup = np.concatenate([L, L[::3]])
    got = ecs.element_consistency(list(dup))
    assert np.abs(got - oracle.element_consistency(dup)).max() < TOL


@needs2
def test_small_jobs_and_per_pair_entries_stay_on_the_primary_device(two_devices):
    import oracle
    from clustassess_b200 import _capi, ecs, synth

    L, _ = synth.config("C1")
    assert np.abs(ecs.element_consistency(list(L)) - oracle.element_consistency(L)).max() < TOL  # 30 partitions: one span
    a, b = synth.config("C2", n=20000)[0]
    assert np.array_equal(ecs.element_sim_elscore(a, b), oracle.disjoint_ecs(a, b))
    dl = _capi.DeviceLabels(L.shape[1], L.shape[0]).upload(L)
    assert not dl.sharded
    dl.close()


@needs2
def test_init_is_refused_while_label_sets_are_alive(two_devices):
    from clustassess_b200 import _capi

    dl = _capi.DeviceLabels(1000, 40)
    with pytest.raises(_capi.ClustAssessError):
        _capi.init([0])
    dl.close()
    assert _capi.init([0]) == 1
