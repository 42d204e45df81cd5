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
    if not files:
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


if __name__ == "__main__":
    main()
