// CPU-only unit test of CommunicationGroup (reference semantics: src/all_to_all_comm.hpp:72-113):
// ranks are cut into grids of `grid_size` consecutive ranks, sampled with spacing `stride`.
// usage: RANK=<r> test_comm_group <grid> <stride>   -> prints "size local_idx member0 member1 ..."
#include <cstdio>
#include <cstdlib>

#include "../host/all_to_all_comm.hpp"

int main(int argc, char** argv)
{
  dj_bootstrap::init(&argc, &argv);
  const int grid = std::atoi(argv[1]), stride = std::atoi(argv[2]);
  CommunicationGroup g(grid, stride);
  std::printf("%d %d", g.size(), g.get_local_idx());
  for (int i = 0; i < g.size(); i++) std::printf(" %d", g.get_global_rank(i));
  std::printf("\n");
  return 0;
}
