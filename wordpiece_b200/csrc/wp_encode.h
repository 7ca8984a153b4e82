// Launch interface of the fused encode kernel (wp_encode.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "wp_table.h"

namespace wp {

struct EncodeParams {
  DeviceVocab vocab;
  const uint8_t *text;              // device, n_bytes of UTF-8
  size_t n_bytes;                   // > 0
  int32_t *ids;                     // device, capacity entries
  unsigned long long capacity;
  uint32_t n_tiles;                 // ceil(n_bytes / tile bytes)
  // scratch, zeroed before every launch:
  unsigned int *ticket;             // tile dispenser
  unsigned long long *tile_state;   // n_tiles look-back words
  unsigned long long *n_ids_out;    // total id count
  unsigned long long *stat_dirty_tiles;
  unsigned long long *stat_long_segments;
};

size_t encode_smem_bytes();
uint32_t encode_tile_bytes();
cudaError_t launch_encode(const EncodeParams &P, cudaStream_t stream);

}  // namespace wp
