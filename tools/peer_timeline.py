"""Timeline of the row-sharded field path from the device time stamps MYFM_PEER_TRACE leaves behind
(csrc/field_sweep.cuh: peer_trace; one record per statistics exchange = per vector of the last field).

    MYFM_PEER_TRACE=gpurun_out/peer torchrun --nproc-per-node 2 bench.py --gpus 2 --steps 10 --warmup 3 --no-parity
    python tools/peer_timeline.py gpurun_out/peer [--skip 200]     -> markdown: median / p90 of every phase, per rank

Phases of one vector of the last field (us):  gather = statistics kernel start -> its last CTA publishes;
gap1 = publish -> draw kernel starts;  wait = draw kernel start -> every peer's sequence number seen;
draw = -> draw kernel done;  gap2 = -> next streaming pass starts;  stream = its block 0 runs;
gap3 = stream done -> next statistics kernel starts;  period = statistics start -> next statistics start.
"""
import glob
import sys

import numpy as np


def main():
    prefix = sys.argv[1]
    skip = int(sys.argv[sys.argv.index("--skip") + 1]) if "--skip" in sys.argv else 200
    files = sorted(glob.glob(prefix + ".rank*.csv"))
    if not files and not glob.glob(prefix + ".phases.rank*.csv"):
        raise SystemExit("no trace files at " + prefix)
    print("| rank | records | gather | gap1 | wait | draw | gap2 | stream | gap3 | period |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for path in files:
        t = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.float64, ndmin=2)
        t = t[np.argsort(t[:, 0])][skip:]
        seq, s0, s1, d0, d1, d2, f0, f1 = t.T
        nxt = np.roll(s0, -1)
        ok = (np.roll(seq, -1) == seq + 1) & (f0 > d2) & (nxt > f1)  # consecutive collectives within one sweep chain
        phases = [s1 - s0, d0 - s1, d1 - d0, d2 - d1, f0 - d2, f1 - f0, nxt - f1, nxt - s0]
        cells = []
        for ph in phases:
            v = ph[ok] / 1e3
            v = v[(v >= 0) & (v < 1e4)]
            cells.append(f"{np.median(v):.1f} / {np.percentile(v, 90):.1f}" if v.size else "-")
        rank = path.rsplit(".rank", 1)[1].split(".")[0]
        print(f"| {rank} | {int(ok.sum())} | " + " | ".join(cells) + " |")
    print("\n(median / 90th percentile, microseconds; gaps include the rest of the sweep where a record is the last of its sweep)")
    phase_files = sorted(glob.glob(prefix + ".phases.rank*.csv"))
    if phase_files:
        names = ["alpha + w0", "lambda_w, mu_w", "w sweep", "lambda_V, mu_V", "V sweep", "merge of owned columns",
                 "e refresh", "to the next sweep's start", "sweep"]
        print("\nPhases of a sweep (median us):\n\n| rank | sweeps | " + " | ".join(names) + " |")
        print("|---|---|" + "---|" * len(names))
        for path in phase_files:
            t = np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.float64, ndmin=2)
            t = t[np.argsort(t[:, 0])][3:]
            d = np.diff(t, axis=1) / 1e3
            nxt = (np.roll(t[:, 0], -1) - t[:, 7])[:-1] / 1e3
            whole = np.diff(t[:, 0]) / 1e3
            ok = whole < 5 * np.median(whole)
            cells = [f"{np.median(d[:, k]):.0f}" for k in range(7)] + [f"{np.median(nxt[ok]):.0f}", f"{np.median(whole[ok]):.0f}"]
            rank = path.rsplit(".rank", 1)[1].split(".")[0]
            print(f"| {rank} | {t.shape[0]} | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    main()
