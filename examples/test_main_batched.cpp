// Batched counterpart of the reference's test_main.cpp (test_main.cpp:12-35): same options, same user command, same
// calls, with a batch dimension.  Instance 0 is the reference's default run; the others get a deterministic offset of
// the initial forward velocity.  Build:
//   g++ -std=c++14 -I include examples/test_main_batched.cpp mhpc_minimal_env_b200/libmhpc_b200.so -Wl,-rpath,$PWD/mhpc_minimal_env_b200 -o test_main_batched
#include <cstdio>
#include <cstdlib>

#include "mhpc_b200.hpp"

int main(int argn, char* argv[]) {
    const int n = argn > 1 ? std::atoi(argv[1]) : 4;
    const bool fixed_budget = argn > 2 && std::atoi(argv[2]) != 0;   // run all max_AL_iter x max_DDP_iter iterations

    // Choose HSDDP options
    HSDDP_OPTION<double> option;
    option.ReB_active = 1;
    option.AL_active = 1;
    option.max_AL_iter = 2;
    option.max_DDP_iter = 3;
    if (fixed_budget) { option.DDP_thresh = -1; option.AL_thresh = -1; }

    // Instantiate MHPCLocomotion
    USRCMD usrcmd;
    usrcmd.vel = 1.5;
    usrcmd.height = 0;
    usrcmd.roll = 0;
    usrcmd.pitch = 0;
    usrcmd.yaw = 0;
    MHPCUserParameters mhpcparams;
    mhpcparams.usrcmd = &usrcmd;
    Gait gait;
    BatchedMHPCLocomotion<double> locomotion(n, &mhpcparams, &gait, option);
    for (int i = 1; i < n; ++i) locomotion.problem(i).x0[7] += 0.01 * i;
    locomotion.enable_timing();
    locomotion.initialization();
    locomotion.solve_mhpc();
    for (const TIME_PER_ITERATION& t : locomotion.time_ddp(0))   // the reference's TIME_BENCHMARK table (help.cpp:3-13)
        std::printf("time_ddp  iter %d  n_bws %d  time_bws %.3f ms  n_fit %d  time_fit %.3f ms  time_partial %.3f ms\n", t.DDP_iter, t.n_bws, t.time_bws, t.n_fit,
                    t.time_fit, t.time_partial);
    locomotion.print_debugInfo(0);
    for (int i = 0; i < n; ++i) {
        std::printf("instance %d: cost %.12f  exp. cost change %.6e  terminal constraint violation %.3e\n", i,
                    locomotion._actual_cost[i], locomotion._exp_cost_change[i], locomotion._tconstr_violation[i]);
        for (const mhpc_iter_record& r : locomotion.records(i))
            std::printf("   AL %d DDP %d  bws %d  line-search trials %d  reg %g  eps %g  cost %.9f\n", r.iter_AL, r.iter_DDP, r.n_bws, r.n_fit, r.reg, r.eps, r.cost);
    }
    return 0;
}
