/*
 * flhost.h -- C entry points of libflhost.so: the FHEController veneer (host/FHEController.h) and the Linformer forward
 * (host/linformer.h) for callers that cannot include C++ headers (the Python tests and bench.py).  A C++ caller -- the
 * reference's main.cpp -- uses FHEController.h directly.
 */
#ifndef FLHOST_H
#define FLHOST_H
#include "../../include/fl_ckks.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct flh_controller flh_controller;
typedef void (*flh_checkpoint_fn)(const char* name, const double* slots, int n, int level, void* user);

const char* flh_last_error(void);
flh_controller* flh_new(int device, unsigned long long key_seed);
void flh_free(flh_controller* c);
fl_ctx* flh_native(flh_controller* c);
/* backend options of host/FHEController.h: cache_gb, auto_rotation_keys, batch_rows, hoist_ladders, max_rows_per_batch;
 * packed_keys = 1 generates the rotation keys of the packed (BSGS) linear layers */
int flh_set_option(flh_controller* c, const char* name, double value);
double flh_rotation_key_bytes(flh_controller* c);   /* device memory held by the automorphism keys */
/* FHEController::generate_context(serialize) + generate_bootstrapping_and_rotation_keys (main.cpp:82-85);
 * log_ring = 0 keeps the reference's 2^15 */
int flh_generate(flh_controller* c, int log_ring, const int* rotations, int n_rot, int bootstrap_slots, int serialize);
int flh_load(flh_controller* c, const char* rotation_file, int bootstrap_slots);   /* main.cpp:88-89 */
int flh_info(flh_controller* c, int* circuit_depth, int* num_slots);

/* encoder1 -> pooler -> classifier -> decrypt (main.cpp:105-123).  timings: up to *n_timings (name, seconds) pairs. */
/* dead_work bit 0: issue the operations main.cpp never reads; bit 1: evaluate the E / F projection under encryption (F1);
 * bit 2: the all-token attention circuit of main_2.cpp (F4); bit 3: packed mode -- every linear layer (attention and FFN) on 128
 * rows per ciphertext through BSGS diagonal matrix products (LinformerForward::set_packed); with bit 0 clear only the
 * operations the logits read are evaluated (2 bootstraps per forward) */
int flh_forward(flh_controller* c, const char* weights_dir, const char* input_dir, const char* tokens_dir, int token_limit, int dead_work,
                int classes, double* logits, flh_checkpoint_fn sink, void* user, char* timing_names, int names_cap, double* timing_seconds,
                int* n_timings, int* tokens);
/* The same for `samples` samples in ONE pass (packed mode, flags bit 3): every ciphertext of the forward carries one element per
   sample, so each kernel launch works on all of them (BASELINE config 5: a batch of samples, ciphertext-parallel).  The samples share
   the weights folder and must have the same number of rows; logits is [samples][classes]. */
int flh_forward_many(flh_controller* c, const char* weights_dir, const char* const* input_dirs, const char* const* tokens_dirs, int samples,
                     int token_limit, int flags, int classes, double* logits, int* tokens);

/* Generic method call for the layout tests: cts / pts are C-ABI handles (borrowed), results are new handles (owned by the caller).
 * method is the FHEController method name, ints / reals its scalar arguments in declaration order. */
int flh_invoke(flh_controller* c, const char* method, fl_elem* const* cts, int n_cts, fl_elem* const* pts, int n_pts, const int* ints, int n_ints,
               const double* reals, int n_reals, fl_elem** out, int out_cap, int* n_out);

#ifdef __cplusplus
}
#endif
#endif
