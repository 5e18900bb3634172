// C wrapper around the tile scheduler of the block-task kernel (feddlib_b200/csrc/tasks.cuh) -- test infrastructure:
// tests/test_task_schedule.py checks the schedule's invariants on the CPU.
#include "../../feddlib_b200/csrc/tasks.cuh"

extern "C" int fb_schedule_tile(int n_nodes, const int *len, const int *ninc, const int64_t *k0, const uint32_t *rec, int rec_words,
                                int ncol, uint64_t *out)
{
    return fb::schedule_tile(n_nodes, len, ninc, k0, rec, rec_words, ncol, out);
}
