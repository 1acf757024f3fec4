// 3-D scalar diffusion, LOD<3,1>: the dim-generic extension BASELINE.json configs 4/5 ask for (the reference
// instantiates LOD<2,1> and LOD<2,2> only, source/LOD.cc:1470-1471).
#include "../LOD.h"

int main(int argc, char *argv[]) {
  using namespace slodhost;
  return run_main<DiffusionProblem<3, 1>, LODParameters<3, 1>>(argc, argv);
}
