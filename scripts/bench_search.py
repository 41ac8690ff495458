"""Private-ANN queries/s: private graph search (graphann.SearchKNN over PIRGraphInfo) on synthetic data of the SIFT
or MS-MARCO shape, B200 path vs the CPU oracle on the same host.  Parameters follow run-private-search.sh:16-18
(step 20, parallel 3) and reproduction/msmarco/reproduce.sh:226-230."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def gen(n, dim, m, shape, seed):
    rng = np.random.default_rng(seed)
    if shape == "sift":      # integers 0..255 stored as f32 (graphann/loader.go:47-51)
        vec = rng.integers(0, 256, (n, dim), dtype=np.uint8).astype(np.float32)
    else:                    # MS-MARCO-shaped: per-dimension sigma decaying 0.82 -> 0.29
        vec = rng.standard_normal((n, dim), dtype=np.float32) * np.linspace(0.82, 0.29, dim, dtype=np.float32)
    graph = rng.integers(0, n, (n, m), dtype=np.int32)          # genRandomGraph (private-search.go:54-69)
    self_loop = graph == np.arange(n, dtype=np.int32)[:, None]
    graph[self_loop] = (graph[self_loop] + 1) % n
    return vec, graph


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="sift", choices=["sift", "msmarco"])
    ap.add_argument("--n", type=int, default=0)
    ap.add_argument("--q", type=int, default=200)
    ap.add_argument("--cpu-q", type=int, default=20)
    ap.add_argument("--k", type=int, default=0)
    ap.add_argument("--step", type=int, default=20)
    ap.add_argument("--parallel", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--clients", type=int, default=1, help="concurrent clients (one per user) sharing the resident DB, one host thread each")
    ap.add_argument("--lanes", type=int, default=1, help="clients of one lock-step group (graphann.SearchKNNLockstep)")
    ap.add_argument("--lane-q", type=int, default=4, help="queries per lane in the lock-step measurement")
    ap.add_argument("--groups", type=int, default=1, help="lock-step groups, one host thread each (host work of one overlaps the GPU work of another)")
    a = ap.parse_args()
    n = a.n or (1000000 if a.shape == "sift" else 3201821)
    dim = 128 if a.shape == "sift" else 192
    k = a.k or (10 if a.shape == "sift" else 100)
    m = 32
    vec, graph = gen(n, dim, m, a.shape, 1)
    queries = vec[np.random.default_rng(2).integers(0, n, a.q)] + np.float32(0.5)

    from pacmann_b200 import cabi, graphann
    from pacmann_b200.keys import mix64
    seed = 7
    f = graphann.GraphANNFrontend(vec, graph, private=True, seed=seed, group_lanes=max(1, a.lanes))
    t0 = time.perf_counter()
    f.Preprocess()
    prep_s = time.perf_counter() - t0
    pir = f.PIR
    wq = vec[np.random.default_rng(3).integers(0, n, 2)] + np.float32(0.5)
    f.SearchKNNBatch(wq, k, a.step, a.parallel)          # warm-up on queries of its own (a repeated query is served from the local cache)
    l0, s0 = cabi.launch_count(), pir.serverQueries
    t0 = time.perf_counter()
    ret, _ = f.SearchKNNBatch(queries, k, a.step, a.parallel)
    dt = time.perf_counter() - t0
    res = dict(shape=a.shape, n=n, dim=dim, m=m, k=k, step=a.step, parallel=a.parallel, queries=a.q,
               gpu_prep_s=prep_s, gpu_pir_prep_s=pir.PreprocessingTime(), gpu_s_per_query=dt / a.q, gpu_qps=a.q / dt,
               gpu_launches=cabi.launch_count() - l0, server_subqueries=pir.serverQueries - s0,
               success_rate=f.succQueryNum / max(1, f.totalQueryNum))
    if a.clients > 1:
        import threading
        fs = [f] + [graphann.GraphANNFrontend(vec, graph, seed=seed + 100 + i, share_db_with=f) for i in range(a.clients - 1)]
        for g in fs[1:]:
            g.Preprocess()
            g.SearchKNNBatch(wq[:1], k, a.step, a.parallel)
        per = max(1, a.q // 2)
        qs = [vec[np.random.default_rng(50 + i).integers(0, n, per)] + np.float32(0.5) for i in range(a.clients)]
        outs = [None] * a.clients

        def work(i):
            outs[i] = fs[i].SearchKNNBatch(qs[i], k, a.step, a.parallel)

        th = [threading.Thread(target=work, args=(i,)) for i in range(a.clients)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        mdt = time.perf_counter() - t0
        res.update(clients=a.clients, multi_client_qps=a.clients * per / mdt, multi_client_s_per_query_per_client=mdt / per)
    if a.lanes > 1:
        import threading
        # one driver thread per group + its pool for the per-lane host work: sized to the cores (oversubscription costs)
        os.environ.setdefault("PM_HOST_THREADS", str(max(1, min(8, (os.cpu_count() or 1) // a.groups))))
        t0 = time.perf_counter()
        groups = []
        for gi in range(a.groups):
            lead = f if gi == 0 else graphann.GraphANNFrontend(vec, graph, seed=seed + 5000 * gi, share_db_with=f, group_lanes=a.lanes)
            if gi:
                lead.Preprocess()
            group = [lead]
            for i in range(1, a.lanes):
                g = graphann.GraphANNFrontend(vec, graph, seed=seed + 5000 * gi + 1000 + i, lane_of=lead, lane=i)
                g.Preprocess()
                group.append(g)
            groups.append(group)
        res["lockstep_group_setup_s"] = time.perf_counter() - t0
        lqs = [vec[np.random.default_rng(70 + gi).integers(0, n, a.lanes * a.lane_q)] + np.float32(0.5) for gi in range(a.groups)]
        # one host thread per group; each warms up (one query per lane) and then waits for the common start
        start, done = threading.Barrier(a.groups + 1), threading.Barrier(a.groups + 1)

        wqs = [vec[np.random.default_rng(170 + gi).integers(0, n, a.lanes)] + np.float32(0.5) for gi in range(a.groups)]

        def drive(gi):
            graphann.SearchKNNLockstep(groups[gi], wqs[gi], k, a.step, a.parallel)
            start.wait()
            graphann.SearchKNNLockstep(groups[gi], lqs[gi], k, a.step, a.parallel)
            done.wait()

        th = [threading.Thread(target=drive, args=(gi,)) for gi in range(a.groups)]
        for t in th:
            t.start()
        start.wait()
        l0 = cabi.launch_count()
        t0 = time.perf_counter()
        done.wait()
        ldt = time.perf_counter() - t0
        for t in th:
            t.join()
        nlq = a.groups * a.lanes * a.lane_q
        res.update(lanes=a.lanes, groups=a.groups, lockstep_queries=nlq, lockstep_qps=nlq / ldt, lockstep_ms_per_step=ldt / (a.lane_q * a.step) * 1e3,
                   lockstep_gpu_launches=cabi.launch_count() - l0)
    if not a.no_cpu:
        from oracle import oracle as o
        raw = o.pack_db(vec, graph)
        o_pir = o.SimpleBatchPianoPIR(n, (dim + m) * 4, m, raw, 8)
        t0 = time.perf_counter()
        o_pir.preprocessing(key_seed=mix64(seed, 1), repl_seed=mix64(seed, 2), threads=os.cpu_count())
        res["cpu_prep_s_all_threads"] = time.perf_counter() - t0
        start = f.StartVertexIds()
        o.search_knn_private(o_pir, vec, graph, start, queries[:2], k, a.step, a.parallel)
        t0 = time.perf_counter()
        o_ret, _, _ = o.search_knn_private(o_pir, vec, graph, start, queries[2:2 + a.cpu_q], k, a.step, a.parallel)
        cdt = time.perf_counter() - t0
        res.update(cpu_s_per_query=cdt / a.cpu_q, cpu_qps=a.cpu_q / cdt, cpu_threads_online=1,
                   reference_published_s_per_query="0.0559 / 0.0640 (SIFT1M, private-search-report.txt:19,44)")
    print(json.dumps(res))


if __name__ == "__main__":
    main()
