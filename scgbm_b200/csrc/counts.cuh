// counts.cuh -- device representation of the count matrix Y and its integer stats
// (K1; replaces max(Y), colSums, rowSums per batch of R/fit.R:94-96,102-105 and the
// input checks of R/fit.R:287-299).
#pragma once
#include "common.cuh"

namespace scgbm {

struct Counts {
  int storage = SCGBM_STORE_U16;  // SCGBM_STORE_U8 / U16 / F32
  int64_t I = 0, J = 0, ld = 0;   // ld = round_up(I, kRowPad); pad rows are zero
  void* data = nullptr;           // device, column-major ld x J
  DevBuf<uint8_t> owned;          // backing store unless aliased
  DevBuf<double> colsum;          // J
  DevBuf<double> rowsum;          // ld x nbatch (this rank's cells only)
  DevBuf<int32_t> batch;          // J zero-based codes (empty when nbatch == 1)
  int nbatch = 1;
  double maxY = 0;                // this rank
  int elem_bytes() const { return storage == SCGBM_STORE_U8 ? 1 : (storage == SCGBM_STORE_U16 ? 2 : 4); }
};

// Upload/pack Y (host or device pointer, any scgbm_dtype) into `c`, computing
// column sums, per-batch row sums and max.  Throws SCGBM_ERR_INVALID on NaN /
// negative entries (R/fit.R:289-294).  storage_req == AUTO picks u16 and falls back
// to f32 if a count does not fit or is non-integral.
void load_counts(Ctx& ctx, const void* Y, int y_dtype, int y_on_device, int64_t I, int64_t J,
                 int64_t ldY, const int32_t* batch_host, int nbatch, int storage_req, Counts& c);

// the same from a column-compressed sparse matrix (host arrays; 0-based row indices, J + 1 pointers)
void load_counts_csc(Ctx& ctx, const int32_t* csc_i, const int64_t* csc_p, const void* csc_x, int x_dtype,
                     int64_t I, int64_t J, const int32_t* batch_host, int nbatch, int storage_req, Counts& c);

// AUTO storage, second step: max(Y) (over all ranks) turned out to be <= 255 -- repack the u16 matrix as u8 (s_Y = 1)
void narrow_counts_to_u8(Ctx& ctx, Counts& c);

// gather rows `ixs` of Y (all columns) into a packed Counts (projection path, R/fit.R:237)
// full_rowsum_dev (I_full doubles, device, optional): row sums of ALL genes of Y (R/fit.R:217), from the same upload
void load_counts_rows(Ctx& ctx, const void* Y, int y_dtype, int y_on_device, int64_t I_full,
                      int64_t J, int64_t ldY, const int64_t* ixs_host, int64_t Ik, Counts& c,
                      double* full_rowsum_dev = nullptr);

}  // namespace scgbm
