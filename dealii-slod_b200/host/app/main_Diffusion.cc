// Mirror of app/main_Diffusion.cc of the reference: 2-D scalar diffusion, LOD<2,1>.
#include "../LOD.h"

int main(int argc, char *argv[]) {
  using namespace slodhost;
  return run_main<DiffusionProblem<2, 1>, LODParameters<2, 1>>(argc, argv);
}
