"""Several GPUs in one process (kgw_options.devices): the work list is cut into cost-balanced shards, one ordinary plan
per device, every device writes its own rows of the caller's matrix.  The keyed generator makes the result independent
of the split, so the sharded call must return exactly the single-device matrix.  On a one-GPU box the shards share
device 0 (the same ordinal listed several times); `gpurun --gpus N` boxes use distinct devices."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
kg = pytest.importorskip("knockoffgwas_b200")


def _devices(n):
    have = kg.device_count()
    return tuple(d % have for d in range(n))


@pytest.fixture(scope="module")
def mixed():
    from knockoffgwas_b200 import synth
    sp = synth.make_problem_fast(1500, 1200, 20, seed=61, n_founders=80, mean_group=2.0)
    synth.add_ibd_pairs(sp, 20)
    synth.add_ibd_families(sp, 3, size=4, segments=4)
    related = np.concatenate([m for m, _ in sp.families])
    unrelated = np.setdiff1d(np.arange(sp.N, dtype=np.int32), related).astype(np.int32)
    return sp, unrelated


def _generate(sp, unrelated, devices):
    return kg.generate(sp.H, sp.p, sp.K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, unrelated=unrelated, ref_global=sp.ref,
                       seed=sp.seed, families=sp.families, options=kg.Options(devices=devices))[0]


@pytest.mark.parametrize("n", [2, 3, 8])
def test_sharded_generate_equals_single_device(mixed, n):
    sp, unrelated = mixed
    one = _generate(sp, unrelated, ())
    many = _generate(sp, unrelated, _devices(n))
    assert np.array_equal(one, many)
    assert one.any()


def test_sharded_plan_partitions_and_bed(mixed):
    """Plan reuse across partitions on a multi-device plan, and the .bed body from it (the rows of the other devices are
    gathered to the first device over peer copies before the transpose)."""
    from knockoffgwas_b200 import synth
    sp, unrelated = mixed
    kw = dict(unrelated=unrelated, ref_global=sp.ref, seed=sp.seed, families=sp.families)
    single = kg.Plan(sp.H, sp.p, sp.K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, **kw)
    multi = kg.Plan(sp.H, sp.p, sp.K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, options=kg.Options(devices=_devices(3)), **kw)
    grng = np.random.default_rng(3)
    for mean_group in (1.0, 6.0):
        groups = synth.make_groups(sp.p, mean_group, grng)
        for pl in (single, multi):
            pl.set_groups(groups)
            pl.run()
        a, _ = single.download()
        b, _ = multi.download()
        assert np.array_equal(a, b), f"mean group {mean_group}"
        bed_a, _ = single.encode_bed()
        bed_b, _ = multi.encode_bed()
        assert np.array_equal(bed_a, bed_b)
    assert multi.info()["launches_per_run"] >= single.info()["launches_per_run"]
    single.close()
    multi.close()


def test_sharded_errors(mixed):
    sp, unrelated = mixed
    with pytest.raises(kg.KgwError):
        _generate(sp, unrelated, (0, 99))
    with pytest.raises(kg.KgwError, match="per device"):
        kg.generate(sp.H, sp.p, sp.K, sp.ref[:, None, :], sp.b, sp.mu, sp.groups, unrelated=unrelated, ref_global=sp.ref,
                    seed=sp.seed, families=sp.families, options=kg.Options(devices=(0, 0)), debug_paths=True)
