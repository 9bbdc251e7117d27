/*
 * libppn_decode — benchmark and profiling hooks.  NOT part of the drop-in interface (include/ppn_decode.h): nothing a
 * host of the parser needs is declared here.  bench.py, scripts/ and the tuning-sweep tests use them.
 *
 *   - ppn_tune writes one process-wide table under a lock; every entry point of ppn_decode.h takes a snapshot of the
 *     table when it starts, so a concurrent writer never changes a call in flight.
 *   - ppn_profile_* is single-threaded by contract (one benchmark thread).
 */
#ifndef PPN_DECODE_BENCH_H_
#define PPN_DECODE_BENCH_H_

#include "ppn_decode.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Per-stage timing of ppn_parse for benchmarks.  After ppn_profile_enable(1) every ppn_parse
 * call (up to 4096) records CUDA events on its stream at the stage boundaries;
 * ppn_profile_read() waits for them and returns the summed milliseconds of the four stages
 * {limb arg-max, decode, NMS, tree parse} and the number of calls covered, then resets. */
int ppn_profile_enable(int32_t on);
int ppn_profile_read(float* stage_ms /*[4]*/, int32_t* n_calls);

/* Device timeline of ppn_parse's kernels: while `dev_records` (device memory, [max_records][4] uint64, initialised
 * by the caller to {~0, 0, ~0, 0}) is set, every kernel a ppn_parse call launches takes the next record and writes
 * {first CTA start, last CTA end, first CTA past its dependency wait, -} in %globaltimer nanoseconds — in launch
 * order: (arg-max, fused parse) per sub-batch on the two-kernel chain, (decode+NMS, arg-max, tree parse) on the
 * three-kernel chain.  NULL switches it off.  One benchmark thread; scripts/timeline.py prints the overlap. */
int ppn_timeline(void* dev_records, int32_t max_records);

/* How the ring arg-max kernel cuts a batch into work items on a GPU of `sms` SMs (host arithmetic only, no launch):
 * info[4] = {matrices per full item, full items, matrices per tail item, items}; first/size (each [max_items] or
 * NULL) receive every item's first matrix and matrix count.  For the test that the items tile [0, B*E) exactly. */
int ppn_debug_argmax_items(const PPNShape* shape, int32_t sms, int32_t* info /*[4]*/, int32_t* first, int32_t* size, int32_t max_items);

/* Benchmark knobs.  key: "argmax.variant" (0 = TMA bulk-copy ring, 1 = direct 128-bit loads),
 * "argmax.stage_bytes", "argmax.stages", "argmax.threads", "argmax.ctas_per_sm",
 * "argmax.split" (-1 auto, 0 thread groups split rows, 1 thread groups take one matrix each),
 * "argmax.dynamic" (1 = ticket scheduling), "argmax.tail_opt", "argmax16.threads", "argmax16.stage_bytes" (ring shape
 * for 16-bit heads), "argmax.cluster" (tiny batches: -1 auto, 0 never, 2/4/8 = CTAs per matrix), "argmax.smem_cap" (stand-alone arg-max: bytes of shared
 * memory the ring may use, 0 = all),
 * "parse.fused" (-1 auto, 0 three-kernel chain, 1 two-kernel chain whenever supported, cutting large batches),
 * "parse.chain_calls", "parse.persist" (three-kernel chain: decode+NMS CTAs per SM of the persistent grids, 0 = one CTA per image), "parse.k12_threads" (decode+NMS CTA size of the three-kernel chain, 0 = auto), "parse.threads", "parse.stage_all" (-1 auto, 0 nothing staged, >= 1 staged whenever it fits),
 * "parse.overlap" (0 serial, 1 decode+NMS on a side stream, 2 single-stream PDL chain = default), "host.chunk_images",
 * "nms.blockwise" (1 = the NMS's round-2 block-by-block phase and n x n ranking instead of the warp wavefront and the bucket sort:
 * a second implementation for A/B timing and for the parity test that compares the two),
 * "head.subs" (fused head: epilogue warps per TMEM lane quadrant), "head.dry" (fused head probes, results invalid: 1 epilogue
 * skipped, 2 an eighth of the MMAs), "head.acc" (16-bit path: 128 = four TMEM accumulators of 128 channels, default two of 256), "timeline.phase" (which in-kernel phase mark ppn_timeline records),
 * "encode.sweep" (1 = address-ordered persistent sweep), "encode.ctas_per_sm".  Returns PPN_E_BADARG for an unknown key. */
int ppn_tune(const char* key, int32_t value);
int ppn_tune_get(const char* key, int32_t* value);

#ifdef __cplusplus
}
#endif
#endif /* PPN_DECODE_BENCH_H_ */
