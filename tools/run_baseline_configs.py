"""Runs BASELINE.json's configurations 1-3 through the reference-facing CLI (`baghera-tool gene-heritability`) on synthetic
annotated-SNP tables of the named shapes and records what the sampler did (profiles/r3_baseline_configs_cli.json).

    python tools/run_baseline_configs.py [cfg1 cfg2 cfg3]

cfg1: 20k SNPs x 200 genes, 4 chains, sweeps 1000 burn-in 250;  cfg2: 1.1M x 15k, 4 chains, sweeps 10000 burn-in 2500
(ref: README.md:64); cfg3: the same with 64 chains (K6 streaming statistics + fp32 trace: the fp64 trace would be 77 GB)."""
import json, os, sys, tempfile, time
import numpy as np, pandas as pd, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import gene_regression as gr
from baghera_b200 import synth
from baghera_b200.cli import main

CFG = {"cfg1": (20_000, 200, 4, 1000, 250), "cfg2": (1_100_000, 15_000, 4, 10000, 2500), "cfg3": (1_100_000, 15_000, 64, 10000, 2500)}
which = sys.argv[1:] or ["cfg1", "cfg2", "cfg3"]
out = {}
tmp = tempfile.mkdtemp()
tables = {}
for name in which:
    M, G, C, sweeps, burn = CFG[name]
    if (M, G) not in tables:
        t0 = time.perf_counter()
        df, truth = synth.make_snp_table(M, G, seed=20211)
        path = os.path.join(tmp, "snps_%d.csv" % M)
        df.to_csv(path, index=False)
        tables[(M, G)] = (path, truth, time.perf_counter() - t0)
    path, truth, t_gen = tables[(M, G)]
    files = [os.path.join(tmp, "%s_%s" % (name, x)) for x in ("genes.csv", "summary.csv", "log.txt")]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rc = main(["gene-heritability", path, *files, "--sweeps", str(sweeps), "-b", str(burn), "--n-chains", str(C), "-m", "normal", "--seed", "7"])
    wall = time.perf_counter() - t0
    assert rc == 0
    lr = gr.analyse_normal.last_run
    st = lr["stats"]
    post = st[burn:]
    su = pd.read_csv(files[1], index_col=0)
    res = pd.read_csv(files[0])
    secs = lr["seconds_total"]
    out[name] = {
        "shape": {"M": M, "G": G, "chains": C, "sweeps": sweeps, "burnin": burn},
        "command": "baghera-tool gene-heritability <snps.csv> genes.csv summary.csv log.txt --sweeps %d -b %d --n-chains %d -m normal" % (sweeps, burn, C),
        "wall_seconds_cli": wall, "sampler_seconds": secs, "ticks": lr["ticks"], "us_per_tick": secs / max(lr["ticks"], 1) * 1e6,
        "chain_leapfrogs": float(st[:, 2].sum()), "chain_leapfrogs_per_sec": float(st[:, 2].sum()) / secs,
        "mean_tree_accept_post_tune": float(post[:, 1].mean()), "mean_tree_accept_tune_last_500": float(st[max(burn - 500, 0):burn, 1].mean()),
        "mean_depth_post_tune": float(post[:, 0].mean()), "leapfrogs_per_draw_post_tune": float(post[:, 2].mean()),
        "divergences_post_tune": int(post[:, 3].sum()), "step_size_post_tune_mean": float(post[:, 5].mean()),
        "summary": {k: {c: float(su.loc[k, c]) for c in ("mean", "sd", "n_eff", "Rhat")} for k in su.index},
        "ess_min_scalar": float(su["n_eff"].min()), "ess_per_sec_min_scalar": float(su["n_eff"].min()) / (secs * sweeps / (sweeps + burn)),
        "rhat_max": float(su["Rhat"].max()),
        "truth": {"e": float(truth["e"]), "mi": float(truth["mi"])},
        "genes_rows": int(len(res)), "stream_stats": bool(lr["stream_stats"]), "trace_dtype": lr["trace_dtype"],
        "peak_hbm_gb": lr["peak_hbm_bytes"] / 1e9, "kernel_launches": int(lr["kernel_launches"]),
    }
    print(name, json.dumps(out[name]), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/r3_baseline_configs.json", "w"), indent=1)
