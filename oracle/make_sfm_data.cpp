// make_sfm_data.cpp — TEST INFRASTRUCTURE ONLY.  Writes an openMVG sfm_data.json (views + one shared pinhole
// intrinsic) with the reference's own sfm::Save, so tests/test_integration_gpu.py can feed the stock and the patched
// openMVG_main_ComputeMatches the files they expect.   usage: make_sfm_data <out.json> <n_views>
#include "openMVG/cameras/cameras.hpp"
#include "openMVG/sfm/sfm_data.hpp"
#include "openMVG/sfm/sfm_data_io.hpp"

#include <cstdio>
#include <cstdlib>
#include <memory>

using namespace openMVG;
using namespace openMVG::sfm;

int main(int argc, char ** argv)
{
  if (argc < 3) { std::fprintf(stderr, "usage: %s out.json n_views\n", argv[0]); return 2; }
  const int n = std::atoi(argv[2]);
  SfM_Data s;
  s.s_root_path = "images";
  s.intrinsics[0] = std::make_shared<cameras::Pinhole_Intrinsic>(1000, 1000, 1000.0, 500.0, 500.0);
  for (int i = 0; i < n; ++i) {
    char name[64]; std::snprintf(name, sizeof name, "img_%04d.jpg", i);
    s.views[i] = std::make_shared<View>(name, i, 0, i, 1000, 1000);
  }
  return Save(s, argv[1], ESfM_Data(VIEWS | INTRINSICS)) ? 0 : 1;
}
