"""Turn the parity report written by the -m gpu tests (tests/parity.py::record) into a markdown table:
    python tests/make_parity_table.py gpurun_out/parity_report.jsonl > profiles/r02_parity.md
Later lines of the same (case, mode) replace earlier ones (the file is appended to run after run)."""
import json
import sys


def main():
    rows = {}
    for line in open(sys.argv[1]):
        line = line.strip()
        if line:
            r = json.loads(line)
            rows[(r["case"], r["mode"])] = r
    print("# Engine against the oracle on BASELINE.json's configs (B200, through the C ABI)\n")
    print("Written by `tests/test_gpu_configs.py` (`-m gpu`). Same ids, style rows, speeds and injected SineGen noise on both sides.")
    print("`pred_dur` and the alignment index map are bit-exact in every row (the tests assert it). Tap columns: max-abs error relative")
    print("to the oracle tap's max-abs. *back half*: the oracle's generator fed the engine's own `gen_in` / `F0` / source spectrum")
    print("(bit-identical inputs; the north star's bar applies here). *end to end*: the engine's waveform against the oracle's,")
    print("nothing substituted - the reference graph itself is chaotic in F0 (an oracle-vs-oracle run under a 1e-6 relative F0")
    print("perturbation gives ~19 dB, `tests/test_oracle.py::test_reference_is_chaotic_in_f0`), so this column is a report, not a bar.\n")
    print("| config | mode | utt | frames | dur abs err (margin to .5) | d | t_en | F0 | N | gen_in | back half: max-abs worst / SNR min (median) dB | end to end: max-abs worst / SNR min (median) dB |")
    print("|---|---|---:|---:|---|---:|---:|---:|---:|---:|---|---|")
    for (case, mode), r in rows.items():
        f = r.get("front")
        fr = (f"{f['duration_abs']:.1e} ({f['duration_margin']:.1e}) | {f['d']:.1e} | {f['t_en']:.1e} | {f['F0']:.1e} | {f['N']:.1e} | {f['gen_in']:.1e}"
              if f else "- | - | - | - | - | -")
        b, e = r["back"], r["e2e"]
        print(f"| {case} | {mode} | {r['utterances']} | {r['frames']} | {fr} | {b['max_abs_worst']:.2e} / {b['snr_min']:.1f} ({b['snr_median']:.1f}) | "
              f"{e['max_abs_worst']:.2e} / {e['snr_min']:.1f} ({e['snr_median']:.1f}) |")


if __name__ == "__main__":
    main()
