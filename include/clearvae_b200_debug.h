/*
 * clearvae_b200_debug.h — test / profiling hooks of libclearvae_b200.so.  NOT part of the drop-in boundary
 * (include/clearvae_b200.h): these entry points carry process-global state or write timelines, so the stateless /
 * re-entrant contract of the public ABI does not apply to them.  Used by tests/ and tools/ only.
 */
#ifndef CLEARVAE_B200_DEBUG_H_
#define CLEARVAE_B200_DEBUG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* test hook: minimum local batch for the tensor-core (tcgen05) latent kernels; default 4096.  Process-global. */
void clearvae_set_latent_tc_min_rows(int32_t rows);

/* profiling hook (tools/conv_timeline.py): when non-NULL, every CTA of clearvae_conv_gemm writes 8 int64
 * %globaltimer stamps (start, prologue done, loads issued, loads landed, accumulator ready, epilogue done, exit) */
int clearvae_debug_conv_timeline(long long* device_buffer);

/* profiling hook (tools/latent_timeline.py): when non-NULL, CTA (0, 0) of the tensor-core latent backward writes clock64 stamps
 * for its first 64 column tiles, [64][16] int64: 0/1 producer, 2-5 MMA issuer, 8-15 epilogue (see csrc/latent_loss.cu) */
int clearvae_debug_latent_timeline(long long* device_buffer);

/* debug: %globaltimer stamps {start, staged, peers ready, pulled} (+2 spare) of CTA 0 for the last 64 peer calls, [64][6] u64 */
int clearvae_peer_timeline(const void* local_base, uint64_t* stamps_host);

#ifdef __cplusplus
}
#endif
#endif /* CLEARVAE_B200_DEBUG_H_ */
