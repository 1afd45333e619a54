// Batched counterpart of the reference's test_main.cpp (test_main.cpp:12-35): same options, same user command, same
// calls, with a batch dimension.  Instance 0 is the reference's default run; the others get a deterministic offset of
// the initial forward velocity.  Build:
//   g++ -std=c++14 -I include examples/test_main_batched.cpp mhpc_minimal_env_b200/libmhpc_b200.so -Wl,-rpath,$PWD/mhpc_minimal_env_b200 -o test_main_batched
#include <cstdio>
#include <cstdlib>

#include "mhpc_b200.hpp"

int main(int argn, char* argv[]) {
    const int n = argn > 1 ? std::atoi(argv[1]) : 4;

    // Choose HSDDP options
    HSDDP_OPTION<double> option;
    option.ReB_active = 1;
    option.AL_active = 1;
    option.max_AL_iter = 2;
    option.max_DDP_iter = 3;

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
    locomotion.initialization();
    locomotion.solve_mhpc();
    locomotion.print_debugInfo(0);
    for (int i = 0; i < n; ++i) {
        std::printf("instance %d: cost %.12f  exp. cost change %.6e  terminal constraint violation %.3e\n", i,
                    locomotion._actual_cost[i], locomotion._exp_cost_change[i], locomotion._tconstr_violation[i]);
        for (const mhpc_iter_record& r : locomotion.records(i))
            std::printf("   AL %d DDP %d  bws %d  line-search trials %d  reg %g  eps %g  cost %.9f\n", r.iter_AL, r.iter_DDP, r.n_bws, r.n_fit, r.reg, r.eps, r.cost);
    }
    return 0;
}
