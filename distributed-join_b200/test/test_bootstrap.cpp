// test_bootstrap.cpp -- CPU check of the MPI-free rendezvous: rank 0 publishes 128 bytes, the other
// ranks of the same launch read exactly those bytes (tests/test_cpp_mirror.py starts the ranks).
#include <cstdio>
#include <cstring>

#include "../host/bootstrap.hpp"

int main(int argc, char** argv)
{
  dj_bootstrap::init(&argc, &argv);
  unsigned char id[128];
  std::memset(id, 0, sizeof(id));
  if (dj_bootstrap::rank() == 0)
    for (int i = 0; i < 128; i++) id[i] = (unsigned char)(i * 7 + (argc > 1 ? argv[1][0] : 1));
  dj_bootstrap::broadcast_from_root(id, sizeof(id), "nccl_id");
  unsigned sum = 0;
  for (int i = 0; i < 128; i++) sum = sum * 31 + id[i];
  std::printf("%d %d %u\n", dj_bootstrap::rank(), dj_bootstrap::size(), sum);
  return 0;
}
