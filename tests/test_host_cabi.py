"""CPU-side checks of the boundary (no GPU, no compute calls): the C-ABI library loads and exports every
symbol ``include/genvc_b200.h`` declares, layout queries work without a device, the weight packer puts
checkpoint tensors where the C side says, argument validation fails loudly, and the replicas plumbing
(shard / broadcast / gather) works across two ``gloo`` processes."""
import ctypes as C
import os
import re
import socket

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "genvc_b200.h")


def _lib():
    from genvc_b200.build import build_library
    from genvc_b200.lib import load_library

    build_library()
    return load_library()


def test_library_exports_every_declared_symbol():
    from genvc_b200.lib import SIGNATURES

    lib = _lib()
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    declared = set(re.findall(r"\b(genvc_[a-z0-9_]+)\s*\(", text))
    assert declared, "no declarations parsed from the header"
    assert declared == set(SIGNATURES), f"header / binding mismatch: {declared ^ set(SIGNATURES)}"
    for name in declared:
        assert getattr(lib, name) is not None


def test_layout_queries_need_no_device():
    from genvc_b200.config import GenVCDims, make_config_dict
    from genvc_b200.weights import blob_layout

    dims = GenVCDims.from_config(make_config_dict(30, 1024, 4))
    total, table = blob_layout(dims)
    names = {t[0] for t in table}
    assert "gpt.h.0.attn.c_attn.weight" in names and "conditioning_perceiver.latents" in names
    assert "mel_head.weight" in names and "final_norm.bias" in names
    # every tensor starts inside the blob, and nothing overlaps
    spans = sorted((off, off + rows * stride) for _, off, rows, cols, stride in table)
    assert spans[-1][1] <= total
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0
    # 377.9 M decode parameters -> the blob is a little larger (embeddings, perceiver, text head)
    assert total * 4 > 1_515_778_056


def test_pack_state_dict_places_tensors_and_rejects_bad_checkpoints():
    from genvc_b200.config import GenVCDims
    from genvc_b200.synth import synth_checkpoint
    from genvc_b200.weights import blob_layout, pack_state_dict

    ck = synth_checkpoint(n_layer=2, d_model=128, n_head=4, seed=5)
    dims = GenVCDims.from_config(ck["config"])
    blob = pack_state_dict(dims, ck["model"])
    _, table = blob_layout(dims)
    for key, off, rows, cols, stride in table:
        t = ck["model"].get("gpt." + key)
        if t is None:
            continue
        got = blob[off: off + rows * stride].view(rows, stride)[:, :cols]
        assert torch.equal(got, t.reshape(rows, cols).float()), key
    bad = dict(ck["model"])
    del bad["gpt.gpt.h.1.mlp.c_fc.weight"]
    with pytest.raises(KeyError):
        pack_state_dict(dims, bad)
    bad = dict(ck["model"])
    bad["gpt.mel_head.weight"] = torch.zeros(7, 3)
    with pytest.raises(ValueError):
        pack_state_dict(dims, bad)


def test_create_rejects_unsupported_shapes_with_a_message():
    from genvc_b200.config import GenVCDims, make_config_dict
    from genvc_b200.lib import GENVC_E_INVALID
    from genvc_b200.weights import c_config

    lib = _lib()
    dims = GenVCDims.from_config(make_config_dict(2, 128, 4))
    cfg = c_config(dims)
    cfg.d_model = 100  # not a multiple of 128
    ctx = C.c_void_p()
    rc = lib.genvc_create(C.byref(cfg), 0, C.byref(ctx))
    assert rc == GENVC_E_INVALID
    assert b"d_model" in lib.genvc_last_error(ctx)
    lib.genvc_destroy(ctx)
    # calls on a context without bound buffers fail with a state error instead of touching the device
    cfg = c_config(dims)
    rc = lib.genvc_create(C.byref(cfg), 0, C.byref(ctx))
    assert rc == 0
    assert lib.genvc_prefill(ctx, None, 1, 47, None) < 0
    assert lib.genvc_last_error(ctx) != b""
    lib.genvc_destroy(ctx)


def test_no_cpu_fallback():
    from genvc_b200.config import GenVCDims, make_config_dict
    from genvc_b200.gpt import GPT

    g = GPT(GenVCDims.from_config(make_config_dict(2, 128, 4)), device="cuda")
    with pytest.raises(RuntimeError):
        g.to("cpu")
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            g.init_gpt_for_inference()


def test_shard_units_partitions_everything():
    from genvc_b200.replicas import shard_units

    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in shard_units(37, r, world))
        assert seen == list(range(37))
    with pytest.raises(ValueError):
        shard_units(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _replica_worker(rank, world, port, q):
    import torch.distributed as dist

    from genvc_b200.config import GenVCDims
    from genvc_b200.replicas import broadcast_blob, gather_ids, shard_units
    from genvc_b200.synth import synth_checkpoint
    from genvc_b200.weights import blob_layout, pack_state_dict

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        ck = synth_checkpoint(n_layer=2, d_model=128, n_head=4, seed=3)
        dims = GenVCDims.from_config(ck["config"])
        n_floats, _ = blob_layout(dims)
        # only rank 0 reads the checkpoint: one broadcast of the packed blob (NCCL on GPUs, gloo here)
        blob = pack_state_dict(dims, ck["model"]) if rank == 0 else None
        got = broadcast_blob(blob, n_floats, rank, world, "cpu")
        ref = pack_state_dict(dims, ck["model"])
        ok_blob = bool(torch.equal(got, ref))
        # units dealt round-robin; ragged per-rank results gathered with padding
        mine = shard_units(5, rank, world)
        local = torch.full((len(mine), 3 + rank), 0, dtype=torch.int64)
        for r, u in enumerate(mine):
            local[r] = u
        parts = gather_ids(local, world, pad=1025)
        ok_gather = all(bool((parts[r] == torch.tensor(shard_units(5, r, world)).view(-1, 1)).all()) and
                        parts[r].shape == (len(shard_units(5, r, world)), 3 + r) for r in range(world))
        q.put((rank, ok_blob, ok_gather))
    finally:
        dist.destroy_process_group()


def test_replicas_two_ranks_gloo():
    import torch.multiprocessing as mp

    _lib()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_replica_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]


@pytest.mark.parametrize("L,D,H", [(30, 1024, 4), (30, 1024, 16), (2, 128, 4), (2, 512, 2)])
def test_decode_stream_and_cache_sizes_follow_the_layout(L, D, H):
    """The decode weight stream holds every unit exactly once: per layer 3D + D + 4D + 4D units and V head units of
    D + 4 floats (stream_layout.h); the KV cache is [L][2][max_batch][H][max_seq][hd].  Layout queries need no device."""
    from genvc_b200.config import GenVCDims, make_config_dict
    from genvc_b200.weights import c_config

    lib = _lib()
    dims = GenVCDims.from_config(make_config_dict(L, D, H))
    cfg = c_config(dims)
    ctx = C.c_void_p()
    assert lib.genvc_create(C.byref(cfg), 0, C.byref(ctx)) == 0
    try:
        grid = lib.genvc_decode_grid(ctx)
        assert grid >= 1
        V = cfg.n_audio_vocab
        expect = (L * 12 * D + V) * (D + 4)
        assert lib.genvc_stream_floats(ctx) == expect
        assert lib.genvc_kv_floats(ctx) == L * 2 * cfg.max_batch * cfg.max_seq * D
        assert lib.genvc_workspace_bytes(ctx) > 0
    finally:
        lib.genvc_destroy(ctx)


def test_plan_batches_and_sharding_cover_the_utterance_set():
    """BASELINE configs[2]/[3] host logic: a fixed utterance set dealt to the ranks (strong scaling), each rank's share
    grouped into batches of at most 8 equal-shape rows; every utterance appears exactly once."""
    from genvc_b200.replicas import plan_batches, shard_units

    for n_utt in (32, 8, 5, 1):
        for world in (1, 2, 4, 8):
            seen = []
            for rank in range(world):
                mine = shard_units(n_utt, rank, world)
                groups = plan_batches(len(mine), 8)
                assert all(1 <= len(gr) <= 8 for gr in groups)
                assert sorted(i for gr in groups for i in gr) == list(range(len(mine)))
                if groups:
                    assert max(len(gr) for gr in groups) - min(len(gr) for gr in groups) <= 1
                seen += [mine[i] for gr in groups for i in gr]
            assert sorted(seen) == list(range(n_utt))
    assert plan_batches(0, 8) == []
    assert [len(g) for g in plan_batches(32, 8)] == [8, 8, 8, 8]
    assert [len(g) for g in plan_batches(9, 8)] == [5, 4]


def test_bench_config_is_identical_in_both_arms():
    """The driver compares the ``config`` objects of the two arms of a workload."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name in ("cfg2", "cfg3", "cfg4"):
        a, b = bench.config_dict(name), bench.config_dict(name)
        assert a == b and set(a) == {"workload", "parallelism", "l2"}
        assert name in a["workload"]
    assert bench.weight_bytes() == 1515778056
    assert bench.kv_bytes(60) == 245760 * 60
