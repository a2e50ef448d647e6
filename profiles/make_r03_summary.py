#!/usr/bin/env python
"""Assemble profiles/r03_summary.md from the raw files of the third session's runs in gpurun_out/ (scratch): python profiles/make_r03_summary.py <final-tag> <keys.txt>
(<keys.txt> = output of profiles/ncu_keys.py over the .ncu-rep files)."""
import json
import re
import subprocess
import sys

T = sys.argv[1] if len(sys.argv) > 1 else "r3k"
KEYS = sys.argv[2] if len(sys.argv) > 2 else "/tmp/r3k_keys.txt"
G = "gpurun_out/"
out = []
A = out.append
last = lambda f: open(f).read().strip().splitlines()[-1]
A(f"# Round 2, third session: evidence (B200, gpurun runs r3_01 .. r3_29 and the final runs r3f / r3g / {T})\n")
A("All timings CUDA events on the launching stream unless a table says ncu (cold-cache, serialised: compare shares).  Raw files: `gpurun_out/r3*` (scratch);")
A("the scripts that produced them: `profiles/run_final_r3.sh`, `profiles/env_timeline.py`, `profiles/train_profile.py`, `profiles/frame_phases.py`, `profiles/shard_emulation.py`, "
  "`profiles/shard_breakdown.py`, `profiles/batching_experiment.py`, `profiles/gather_pieces.py`, `profiles/ncu_keys.py`, `profiles/summarize_launches.py`, `profiles/sass_summary.py`; this file: `profiles/make_r03_summary.py`.\n")
A(f"## 1. Bench lines of the final run ({T}: final defaults of the round)\n")
for name, f in (("python bench.py (N = 1, 800x800 three-pass toaster frame)", f"{G}{T}_bench.json"), ("python bench.py --impl reference --steps 3 --warmup 1", f"{G}{T}_bench_ref.json"),
                ("python bench.py --config neus --steps 3 (BASELINE config 4)", f"{G}{T}_bench_neus.json")):
    A(f"### `{name}`\n\n```json\n" + last(f) + "\n```\n")
tests = [l for l in open(f"{G}{T}_tests.log").read().splitlines() if " passed" in l][-1]
A(f"`pytest tests -m gpu`: {tests.strip()} ({T}_tests.log).  `smoke()` ({T}_smoke.log):\n\n```\n" + "\n".join(l[:220] for l in open(f"{G}{T}_smoke.log").read().strip().splitlines()[-2:]) + "\n```\n")
A(f"## 2. ncu launch list of `bench.py --steps 1 --warmup 3 --no-extra-warmup ...` (whole frames between the first two L2 flushes, {T})\n")
A(open(f"{G}{T}_launches.md").read().strip() + "\n")
d = json.loads(last(f"{G}{T}_bench.json"))
m = re.search(r"k_env_tc[^|]*\|[^|]*\|[^|]*\|\s*([0-9.]+)%", open(f"{G}{T}_launches.md").read())
A(f"`k_env_tc` share under ncu {m.group(1) if m else '?'} %; in the bench line (`roofline.kernel_share_of_step`, CUDA events) {100 * d['roofline']['kernel_share_of_step']:.1f} %.\n")
A(f"## 3. `ncu --set full --clock-control none` captures (profiles/ncu_keys.py; inference kernels {T}, training kernels {T} / r3f)\n\n```\n" + open(KEYS).read().strip() + "\n```\n")
A("""Reading: `k_env_tc` (main-pass launch, ~4.25 M samples): DRAM ~545 MB read + ~498 MB written for 0.95 GB algorithmic (ratio 1.1), tensor-memory pipe ~62 % of
elapsed cycles, `sm__pipe_tensor_cycles_active_realtime` 50.8 % of elapsed (raw page of the r3j capture; 53 % in round 1), issue slots ~37 %, shared-memory bank
conflicts 217.6 M of 559.8 M LSU wavefronts = 38.9 % by the raw counter (loads 168.1 M, stores 49.3 M; 822.8 M tensor-core operand wavefronts go through the same port).  The source page
attributes only 36.7 M excessive wavefronts to SASS instructions, ALL of them the IDE warps' 4-byte operand stores (STS at the layer-0 operand buffers, 4-way: rows 8 apart share a bank); the
epilogue's 16-byte operand stores and bias loads are conflict-free, so what the code controls is 6.5 % of the LSU wavefronts, on warps that have ~4x the time they need.  `k_env_tc<SAVE>` (training forward, 70 k samples = 140 k rows): adds 430 MB of activation stores and 4.5 MB of masks.
`k_chain_tc<0>` (= `k_env_bwd_tc` in the r3f build; training backward of env_net): 253 us, 431 MB written (three [140 k, 256] fp32 gradient tensors + d y + d x0 = 476 MB
algorithmic) at 1.9 TB/s, a 7-tile-per-SM launch with the forward kernel's per-tile latency chain; no shared-memory bank conflicts (5 k of 3.8 M wavefronts).\n""")
A("## 4. k_env_tc: clock64 timeline of CTA 0 and timing experiments\n")
A(f"Final build ({T}, `profiles/env_timeline.py`, M = 56,832 samples = 6 tiles per SM; cycles at 1.965 GHz):\n\n```\n" + "\n".join(l[:400] for l in open(f"{G}{T}_env_timeline.txt").read().strip().splitlines()[-6:]) + "\n```\n")
A("Timing experiments (`ENVIDR_ENV_TC_DEBUG`, run r3_10; whole field forward over 2 M samples, geometry + shading included, results of the experiments are garbage by construction):\n")
A("| variant | ms / 2 M samples | tile period (cycles) |\n|---|---:|---:|\n| baseline (reordered issue, 3 products) | 4.30 | ~23.4 k |\n| no weight copies (producer only signals) | 4.21 | ~23.2 k |\n| no IDE arithmetic | 4.13 | ~22.2 k |\n| hi*hi product only (1/3 of the tensor work) | 3.89 | ~20.2 k |\n| no weights + no IDE | 4.04 | ~22.1 k |\n| all three | 3.63 | ~18.7 k |\n")
A("""=> the tile is a latency chain: with a third of the MMA work, no weight traffic and no IDE arithmetic it still takes 18.7 k cycles (14.2 k cycles of MMA work in the
real kernel).  Layer-2 detail of the timeline: the epilogue publishes a chunk pair (64 columns) every ~900 cycles whether 8 or 16 warps convert (r3_11 / r3_12), and the
issuer needs ~850 cycles per chunk (two K steps behind a ring handshake of ~310-425 cycles each).\n""")
A("Variants measured this session (frame = bench.py --no-cpu-baseline --no-gpu-reference --no-train --no-density --no-sweep, `k_env_tc` ms per frame from CUDA events; boxes differ by ~2 %):\n")
A("| build | k_env_tc ms / frame | frame ms | run |\n|---|---:|---:|---|\n| start of session (round-2 build) | 8.04 | 15.96 | r3_02 (c1) |\n| CTA pair without relay warps (tensor-map TMA signalling the leader, named barrier + one remote arrive), 3 x 16 KB ring | 8.68 (9.39 with relays) | 16.71 | r3_01 |\n| + IDE warps sleeping while an epilogue drains (removed) | 8.04 | 15.96 | r3_02 |\n| + next tile's layer 0 issued ahead of the 16-wide last layer (kept) | 7.61 - 7.91 | 15.47 - 15.90 | r3_03, r3_14, r3_17, r3f |\n| + 6 x 8 KB ring stages, hi / lo halves of a K step separately (removed) | 8.02 | 15.87 | r3_04 |\n| + 16 epilogue warps on 16-column halves, 12 IDE warps (removed) | 7.90 | 15.76 | r3_12 |\n| weight multicast across clusters of 2 / 4 CTAs (opt-in) | same as default / 1.7x slower (2 M-sample forward 4.98 vs 4.94 / 8.43 ms) | | r3_09 |\n| + n_step cap 16 / secondary floor 8 in the logged geometry passes (kept; not a k_env_tc change) | 7.74 - 7.87 | 15.30 - 15.40 | r3_23, r3g |\n| + 2 host synchronisations per frame instead of 6 (kept, neutral) | 7.84 | 15.38 | r3_28 |\n| + records read in place through an index instead of a ray-ordered copy (kept) | 7.78 - 7.91 | 15.19 - 15.33 | r3_29, r3j |\n| + 16-column accumulator read-out with the next tcgen05.ld in flight / no proxy fence (timing experiment) / critical warps on the highest ids (all reverted, neutral) | 7.92 - 7.93 | 15.24 - 15.28 | r3_30, r3_31 |\n| + last layer's epilogue on epilogue group 1, group 0 goes straight to the next tile's layer 0 (kept) | 7.74 - 7.78 | 15.03 - 15.22 | r3_32, r3k |\n")
A("## 5. Frame phases, fixed cost per frame, batching of the logged passes\n")
A("`profiles/frame_phases.py` (CUDA events around the phases of `render.render`, 800x800): before the index form of the log gather (run r3_25) and after (run r3_29):\n\n```\n" + open(f"{G}r3_25_phases.txt").read().strip().splitlines()[-1] + "\n" +
  "frame 15.41 ms | render_rays 3.12, last_stats 0.02, render_rays 1.30, last_stats 0.01, prepare_from_log 0.12, shade_prepared 2.23, prepare_from_log 0.16, shade_prepared 7.67 | rest 0.78 ms\n```\n")
A("`profiles/shard_emulation.py` (run r3_20, before the batching change): one rank's share of the 1600x1600 frame on ONE GPU:\n\n```\n" + open(f"{G}r3_20_shard.txt").read().strip() + "\n```\n")
A("fit: 2.7 ms fixed + 53.1 ms / N.  ncu launch lists of the N = 8 share vs the whole frame (run r3_21): kernel time 8.69 ms of 9.31 ms; against 1/8 of the whole frame's kernel times: `k_geom_tc` +0.63 ms, `k_march_compact` +0.47, `k_composite_compact` +0.34, `k_env_tc` +0.24, `k_composite_replay` +0.15, idle 0.6.  After the batching change (run r3_28): world 8 share 8.99 ms (0.78).\n")
A("Batching experiment (`profiles/batching_experiment.py`, run r3_22; image bit-identical in every row):\n\n```\n" + open(f"{G}r3_22_cap.txt").read().strip() + "\n```\n")
A(f"## 6. Train step (config 3 shape, 4,096 rays, ~70 k samples; `profiles/train_profile.py`, {T})\n")
tp = open(f"{G}{T}_train_profile.txt").read()
m = re.search(r"\{'rays'.*?\}", tp, re.S)
A("```\n" + (m.group(0) if m else "") + "\n```\n")
tl = [l[:60] + l[108:200] for l in tp.splitlines() if ("k_" in l or "Fused" in l or "Self CUDA time total" in l or "eager step" in l)]
A("torch.profiler, eager step, top CUDA entries (name | self CUDA | % | total | avg | calls):\n\n```\n" + "\n".join(tl[:26]) + "\n```\n")
t = d["train_step"]
A(f"`bench.py` key `train_step` of the same run: {t['ms_per_step_fwd_bwd']:.2f} ms with the fused env_net / head kernels vs {t['ms_per_step_fwd_bwd_env_per_layer']:.2f} ms with env_net as one launch per layer and direction, graph replay; start of the session 4.96 ms; the real reference `Trainer.train_step` on its own kernels: {d['gpu_reference']['train']['ms_per_step_fwd_bwd']:.1f} ms.\n")
A("## 7. k_geom_tc: level-0 hash table staged in shared memory (run r3_17)\n")
A("ncu over 12 mid-loop launches (`--launch-skip 120 --launch-count 12`), averages:\n\n| | staged (ENVIDR_GEOM_STAGE=1) | not staged (default) |\n|---|---:|---:|\n| gpu__time_duration (us) | 49.3 | 48.1 |\n| l1tex hit rate % | 70.0 | 71.3 |\n| l1tex throughput % | 20.8 | 21.2 |\n| lts throughput % | 7.2 | 7.3 |\n| issue slots active % | 22.3 | 22.8 |\n| frame ms | 15.90 | 15.78 |\n\nNo gain: the 32 KB table already lives in L1.  Level 1 (23^3 x 8 B = 97 KB) does not fit next to the 150 KB of weight images and operands.\n")
A("## 8. Fitted scene (run r3_26; `fitted_scene` key of the bench line above)\n\n```\n" + open(f"{G}r3_26_fit.txt").read().strip().split("\n", 3)[-1] + "\n```\n")
A("## 9. Multi-GPU runs of the session (gpurun --gpus N; `torchrun ... bench.py --gpus N --steps 6 --warmup 3`)\n")
for name, f in (("N = 2 (run r3h)", f"{G}r3h_bench_n2.json"), ("N = 8 before the gather kernel (run r3h: torch index_select de-interleave, every rank reads the frame back in e2e)", f"{G}r3h_bench_n8.json"),
                ("N = 8 with envidr_gather_rows and the e2e read-back on rank 0 (run r3i)", f"{G}r3i_bench_n8.json"),
                ("N = 4, final build (run r3l)", f"{G}r3l_bench_n4.json")):
    A(f"### {name}\n\n```json\n" + last(f) + "\n```\n")
A("`profiles/shard_breakdown.py` at N = 8 (run r3_27, before the gather kernel): per rank (frame ms, render ms):\n\n```\n" + "\n".join(l for l in open(f"{G}r3_27_shard8.txt").read().splitlines() if l.startswith("world") or l.startswith("   max")) + "\n```\n")
A("`profiles/gather_pieces.py` at N = 2 (`[1.28 M, 8]` fp32 per rank): pack 0.077 ms, all_gather_into_tensor 0.120 ms, torch index_select 1.326 ms, slicing 0.003 ms -> the de-interleave is now `envidr_gather_rows` (one float4 per thread).\n")
sass = subprocess.run([sys.executable, "profiles/sass_summary.py"], capture_output=True, text=True).stdout.strip().splitlines()[-20:]
A("## 10. SASS opcode summary of the final library (profiles/sass_summary.py; cuobjdump -sass, sm_100a)\n\n" + "\n".join(sass) + "\n")
A("UTCHMMA = tcgen05.mma kind::f16 (`.2CTA`: cta_group::2), UTCQMMA = kind::f8f6f4, LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk (the CTA-pair kernel loads its weight halves with the tensor-map form UTMALDG instead), UTCBAR = tcgen05.commit, SYNCS = mbarrier operations.  `k_env_tc<CTAS, F8, MC, SAVE>`: `<1, false, 1, false>` is the inference default, `<1, false, 1, true>` the training forward; `k_chain_tc<0|1|2>` = env_net backward / generic MLP backward / generic MLP forward.\n")
open("profiles/r03_summary.md", "w").write("\n".join(out))
print("written", sum(len(x) for x in out))
