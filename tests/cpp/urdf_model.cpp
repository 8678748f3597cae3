// model_t{urdf_path} of the C++ facade (ddpb.hpp) -- the counterpart of the reference's
// pinocchio::model_t<scalar_t>{"~/pinocchio/models/others/robots/ur_description/urdf/ur5_gripper.urdf"}
// (/root/reference/test/pinocchio_ddp.cpp:30).  Needs no GPU: model construction is host code.
#include <cstdio>
#include <fstream>

#include "ddpb/ddpb.hpp"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  try {
    ddpb::model_t model{std::string(argv[1])};
    std::printf("nq %ld nv %ld frames %ld\n", model.configuration_dim(), model.tangent_dim(), model.n_frames());
    for (ddpb::index_t i = 0; i < model.n_frames(); ++i) std::printf("frame %ld %s\n", i, model.frame_name(i));
    ddpb::model_t moved = std::move(model);
    std::printf("moved nq %ld\n", moved.configuration_dim());
  } catch (ddpb::error const& e) {
    std::printf("%s (code %d)\n", e.what(), e.code);
    return 1;
  }
  return 0;
}
