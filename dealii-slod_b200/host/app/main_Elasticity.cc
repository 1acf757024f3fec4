// Mirror of app/main_Elasticity.cc of the reference: 2-D linear elasticity, LOD<2,2>.
#include "../LOD.h"

int main(int argc, char *argv[]) {
  using namespace slodhost;
  return run_main<ElasticityProblem<2, 2>, LODParameters<2, 2>>(argc, argv);
}
