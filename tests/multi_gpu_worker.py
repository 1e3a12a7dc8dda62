"""Worker for tests/test_gpu_multi.py: run under torchrun, one rank per GPU (NCCL carries the rendezvous and the host-side
metadata).  On a box with fewer GPUs than ranks the ranks SHARE a GPU: NCCL refuses two ranks on one device, so the
rendezvous runs on gloo, and every row still moves the way it does between GPUs -- through the library's peer-memory
exchange (CUDA IPC works between two processes on one device), only time-sliced."""

from __future__ import annotations

import os
import sys
import time
from pathlib import Path

os.environ["TZ"] = "UTC"
os.environ.setdefault("MSC_SCAN_RUNS_MIN_ROWS", "1")  # the streaming aggregate over sorted runs also for this test's small tables
time.tzset()
ROOT = Path(__file__).resolve().parent.parent
for p in (ROOT, ROOT / "tests", ROOT / "bench"):
    sys.path.insert(0, str(p))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import cases  # noqa: E402
import gen_tpch  # noqa: E402
from minispark_b200 import CudaExecutionEngine  # noqa: E402
from oracle import py_oracle as O  # noqa: E402


def main() -> None:
    local_rank = int(os.environ["LOCAL_RANK"])
    shared = torch.cuda.device_count() < int(os.environ["WORLD_SIZE"])
    device = local_rank % torch.cuda.device_count()
    torch.cuda.set_device(device)
    if shared:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", device))
    rank, world = dist.get_rank(), dist.get_world_size()
    folder = Path(sys.argv[1])
    lineitem = folder / "lineitem.bin"
    if rank == 0:
        gen_tpch.write_table(lineitem, "lineitem", sf=0.004, rows_per_block=2048)  # ~24k rows in 12 blocks
    dist.barrier()
    ns = cases.namespace()

    def high_card(engine):  # noqa: ANN001, ANN202
        return ns.DataFrame(engine).table(str(lineitem)).group_by(ns.Col("l_orderkey")).agg(
            ns.F.sum(ns.Col("l_quantity")).alias("q"), ns.F.avg(ns.Col("l_extendedprice")).alias("p"), ns.F.count())

    with CudaExecutionEngine(device=device) as engine:
        assert engine.shard == (rank, world)
        peer = os.environ.get("MINISPARK_PEER_EXCHANGE", "1") != "0"
        # low-cardinality GROUP BY: partial tables merged on every rank, every rank gets the full answer
        got = cases.q1(ns, str(lineitem), engine).collect()
        assert engine.last_stats["exchange"].startswith("nvlink peer stores" if peer else "nccl all-gather"), engine.last_stats
        assert engine.last_stats["result_partitioned"] is False
        O.assert_rows_equal(got, O.run_task(cases.q1(ns, str(lineitem)).task, wire=True), rel=5e-7)
        # the same query prepared: pass 1 merges through the partial-table all-gather, later passes exchange the partial
        # tables over NVLink peer memory inside the specialised scan kernel (msc_dense_fused_peer); every rank must hold
        # the same complete answer
        prepared = engine.prepare(cases.q1(ns, str(lineitem)).task)
        want_f64 = {r["l_returnflag"]: r for r in O.run_task(cases.q1(ns, str(lineitem)).task, wire=False)}
        exchanges = []
        for _ in range(4):
            final, _ms = prepared.run()
            names = [n for n, _ in prepared.plan.schema]
            keys = final.cols[0].dict.export()
            cols = [final.column_numpy(i) for i in range(len(names))]
            rows = {keys[int(cols[0][r])]: {n: cols[i][r].item() for i, n in enumerate(names) if i} for r in range(final.nrows)}
            exchanges.append(prepared.scan_stats.get("exchange", "nccl"))
            engine.release_query()
            assert sorted(rows) == sorted(want_f64), (sorted(rows), sorted(want_f64))
            for k, ref in want_f64.items():
                for name, v in rows[k].items():
                    assert abs(v - ref[name]) <= 1e-9 * max(abs(ref[name]), 1e-300), (k, name, v, ref[name])
        if os.environ.get("MINISPARK_PEER_MERGE", "1") != "0":
            assert exchanges[0] == "nccl" and all(x.startswith("nvlink") for x in exchanges[1:]), exchanges
        # high-cardinality GROUP BY through hash partitioning + the row exchange: ranks hold disjoint sets of keys ...
        os.environ["MSC_EXCHANGE_GATHER_MAX"] = "0"
        rel, schema = engine.execute_to_device(high_card(engine).task)
        # (lineitem is clustered by l_orderkey and sharded in row order: the partial results ascend and the ranks' key ranges
        # follow each other, so only the group that straddles two ranks is merged -- its row travels, nothing else does)
        assert engine.last_stats["exchange"].startswith("boundary rows"), engine.last_stats
        assert engine.last_stats["result_partitioned"] is True and rel.partitioned
        assert engine.last_stats["exchange_rows_sent"] <= 1
        mine = rel.column_numpy(0).tolist()
        boundary_rows = {tuple(rel.column_numpy(i).tolist()) for i in range(len(schema))}
        engine.release_query()
        # the same through the general machinery: a range-partitioned exchange of all partial rows + a streaming final aggregate
        os.environ["MSC_EXCHANGE_BOUNDARY"] = "0"
        rel, schema = engine.execute_to_device(high_card(engine).task)
        os.environ.pop("MSC_EXCHANGE_BOUNDARY")
        assert engine.last_stats["exchange"].startswith("nvlink peer push (range" if peer else "nccl send/recv group (range"), engine.last_stats
        assert engine.last_stats["exchange_rows_sent"] <= 1  # at most the one key that straddles two ranks changes rank
        assert {tuple(rel.column_numpy(i).tolist()) for i in range(len(schema))} == boundary_rows
        engine.release_query()
        gathered: list = [None] * world
        dist.all_gather_object(gathered, mine)
        keys = [k for part in gathered for k in part]
        assert len(keys) == len(set(keys)), "a key was aggregated on two ranks"
        # ... the same with hashed routing (what unsorted keys get): about half of the partial rows change rank
        os.environ["MSC_EXCHANGE_RANGE"] = "0"
        rel, schema = engine.execute_to_device(high_card(engine).task)
        assert "(hash partitioned)" in engine.last_stats["exchange"] and engine.last_stats["exchange_rows_sent"] > rel.nrows // 4
        hashed = rel.column_numpy(0).tolist()
        engine.release_query()
        os.environ.pop("MSC_EXCHANGE_RANGE")
        dist.all_gather_object(gathered, hashed)
        assert sorted(k for part in gathered for k in part) == sorted(keys)
        # ... and collect() gathers them: every rank returns the complete result
        full = high_card(engine).collect()
        assert engine.last_stats["result_partitioned"] is False
        os.environ.pop("MSC_EXCHANGE_GATHER_MAX")
        assert sorted(r["l_orderkey"] for r in full) == sorted(keys)
        O.assert_rows_equal(full, O.run_task(high_card(None).task, wire=True), rel=5e-7)
        # the same through the small-table path (every partial row to every rank)
        O.assert_rows_equal(high_card(engine).collect(), O.run_task(high_card(None).task, wire=True), rel=5e-7)
        # filter / project scans run rank-local; collect() gathers the parts in rank order = the table's row order
        def scan(e):  # noqa: ANN001, ANN202
            return (ns.DataFrame(e).table(str(lineitem)).filter(ns.Col("l_quantity") > 49)
                    .select(ns.Col("l_orderkey"), ns.Col("l_linenumber"), ns.Col("l_shipmode"), (ns.Col("l_extendedprice") * 2).alias("p2")))

        rows = scan(engine).collect()
        O.assert_rows_equal(rows, O.run_task(scan(None).task, wire=True), ordered=True)
        rel, _ = engine.execute_to_device(scan(engine).task)
        assert rel.partitioned and 0 < rel.nrows < len(rows)
        engine.release_query()
        # JOIN: both sides are co-partitioned on the key (msc_partition + all-to-all), joined locally, and the
        # aggregate above merges the per-rank partials; string columns travel as codes of rank-independent dictionaries
        orders = folder / "orders.bin"
        if rank == 0:
            gen_tpch.write_table(orders, "orders", sf=0.004, rows_per_block=1024)
        dist.barrier()

        def join_agg(e):  # noqa: ANN001, ANN202
            o = ns.DataFrame(e).table(str(orders)).alias("o")
            l = ns.DataFrame().table(str(lineitem)).alias("l")
            return (o.join(l, on=ns.Col("o.o_orderkey") == ns.Col("l.l_orderkey"), how="inner")
                    .filter(ns.Col("o.o_orderdate").between("1994-01-01", "1995-12-31"))
                    .filter(ns.Col("l.l_shipmode").like("%AIR%"))
                    .group_by(ns.Col("o.o_orderpriority")).agg(ns.F.count(), ns.F.sum(ns.Col("l.l_extendedprice")).alias("rev")))

        got = join_agg(engine).collect()
        want = O.run_task(join_agg(None).task, wire=True)
        assert len(want) > 0
        O.assert_rows_equal(got, want, rel=5e-7)

        def join_rows(e):  # noqa: ANN001, ANN202
            o = ns.DataFrame(e).table(str(orders)).alias("o")
            l = ns.DataFrame().table(str(lineitem)).alias("l")
            return (o.join(l, on=ns.Col("o.o_orderkey") == ns.Col("l.l_orderkey"), how="inner")
                    .filter(ns.Col("l.l_quantity") > 48)
                    .select(ns.Col("o.o_orderkey"), ns.Col("o.o_orderstatus"), ns.Col("l.l_linenumber"), ns.Col("l.l_shipmode")))

        rows = join_rows(engine).collect()
        O.assert_rows_equal(rows, O.run_task(join_rows(None).task, wire=True))
        rel, _ = engine.execute_to_device(join_rows(engine).task)  # without the gather every rank holds the pairs of its keys
        assert rel.partitioned
        part_rows = [None] * world
        dist.all_gather_object(part_rows, rel.nrows)
        engine.release_query()
        assert sum(part_rows) == len(rows) and max(part_rows) < len(rows)

        def agg_then_join(e):  # noqa: ANN001, ANN202  -- a REPLICATED relation (a merged GROUP BY) as a join side: it must enter once
            per_order = (ns.DataFrame(e).table(str(lineitem)).group_by(ns.Col("l_shipmode")).agg(ns.F.count().alias("n"))
                         .select(ns.Col("l_shipmode").alias("mode2"), ns.Col("n")))
            m = ns.DataFrame().table(str(folder / "modes.bin"))
            return per_order.join(m, on=ns.Col("mode2") == ns.Col("mode"), how="inner")

        def join_str_key(e):  # noqa: ANN001, ANN202  -- string join key: lineitem's ship modes against a 3-row lookup table
            m = ns.DataFrame(e).table(str(folder / "modes.bin")).alias("m")
            l = ns.DataFrame().table(str(lineitem)).alias("l")
            return (m.join(l, on=ns.Col("m.mode") == ns.Col("l.l_shipmode"), how="inner")
                    .group_by(ns.Col("m.label")).agg(ns.F.count(), ns.F.sum(ns.Col("l.l_quantity")).alias("q")))

        if rank == 0:
            ns.BlockFile(folder / "modes.bin").write_rows([{"mode": "AIR", "label": "sky"}, {"mode": "REG AIR", "label": "sky"},
                                                           {"mode": "SHIP", "label": "sea"}, {"mode": "NOPE", "label": "none"}])
        dist.barrier()
        got = join_str_key(engine).collect()
        O.assert_rows_equal(got, O.run_task(join_str_key(None).task, wire=True), rel=5e-7)
        got = agg_then_join(engine).collect()
        O.assert_rows_equal(got, O.run_task(agg_then_join(None).task, wire=True))
    dist.barrier()
    if rank == 0:
        print("multi-gpu ok", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
