#include "bootstrap.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "communicator.hpp"
#include "error.hpp"

namespace dj_bootstrap {

static int g_rank = 0, g_size = 1;
static Communicator* g_comm = nullptr;

static int env_int(const char* a, const char* b, const char* c, int dflt)
{
  for (const char* name : {a, b, c}) {
    if (!name) continue;
    const char* v = std::getenv(name);
    if (v && *v) return std::atoi(v);
  }
  return dflt;
}

void init(int*, char***)
{
  g_rank = env_int("RANK", "OMPI_COMM_WORLD_RANK", "PMI_RANK", 0);
  g_size = env_int("WORLD_SIZE", "OMPI_COMM_WORLD_SIZE", "PMI_SIZE", 1);
}

void finalize() { g_comm = nullptr; }
int rank() { return g_rank; }
int size() { return g_size; }

double wtime()
{
  using clock = std::chrono::steady_clock;
  return std::chrono::duration<double>(clock::now().time_since_epoch()).count();
}

static std::string rendezvous_path(const char* tag)
{
  const char* dir  = std::getenv("DJ_RENDEZVOUS_DIR");
  const char* port = std::getenv("MASTER_PORT");
  const char* job  = std::getenv("DJ_JOB_ID");
  std::string p    = dir ? dir : "/tmp";
  p += "/dj_b200_";
  p += job ? job : (port ? port : "default");
  // every rank of one launch has the same parent (the launcher's agent): a per-launch nonce, so a
  // file left behind by a killed run on the same port can never be mistaken for this run's
  p += "_" + std::to_string((long)::getppid()) + "_";
  p += tag;
  return p;
}

void broadcast_from_root(void* buf, std::size_t bytes, const char* tag)
{
  if (g_size == 1) return;
  const std::string path = rendezvous_path(tag), tmp = path + ".tmp";
  if (g_rank == 0) {
    // fresh private file (never follows a planted symlink), published atomically by rename
    ::unlink(path.c_str());
    ::unlink(tmp.c_str());
    const int fd = ::open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL | O_NOFOLLOW, 0600);
    CHECK_ERROR(fd >= 0, true, "open rendezvous file");
    CHECK_ERROR(::write(fd, buf, bytes) == (ssize_t)bytes, true, "write rendezvous file");
    ::close(fd);
    CHECK_ERROR(std::rename(tmp.c_str(), path.c_str()), 0, "publish rendezvous file");
  } else {
    for (int tries = 0;; tries++) {
      const int fd = ::open(path.c_str(), O_RDONLY | O_NOFOLLOW);
      if (fd >= 0) {
        struct stat sb;
        const bool mine = ::fstat(fd, &sb) == 0 && sb.st_uid == ::getuid() && S_ISREG(sb.st_mode);
        const ssize_t got = mine ? ::read(fd, buf, bytes) : -1;
        ::close(fd);
        if (got == (ssize_t)bytes) break;
      }
      if (tries > 60000) {
        std::fprintf(stderr, "ERROR: rank %d timed out waiting for %s\n", g_rank, path.c_str());
        std::exit(1);
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(1));
    }
  }
}

void set_communicator(Communicator* c) { g_comm = c; }

static std::vector<int64_t> gather_one(int64_t v)
{
  std::vector<int64_t> all(g_size, v);
  if (g_size > 1) {
    CHECK_ERROR(g_comm != nullptr, true, "collective before the communicator exists");
    g_comm->allgather_i64(&v, 1, all.data());
  }
  return all;
}

void barrier()
{
  if (g_size > 1) gather_one(0);
  // rank 0 removes the rendezvous file once everybody has passed the first barrier
  static bool cleaned = false;
  if (!cleaned && g_rank == 0 && g_size > 1) {
    ::unlink(rendezvous_path("nccl_id").c_str());
    cleaned = true;
  }
}

double allreduce_max(double v)
{
  int64_t bits;
  std::memcpy(&bits, &v, 8);
  double m = v;
  for (int64_t b : gather_one(bits)) {
    double x;
    std::memcpy(&x, &b, 8);
    if (x > m) m = x;
  }
  return m;
}

int64_t allreduce_sum(int64_t v)
{
  int64_t s = 0;
  for (int64_t x : gather_one(v)) s += x;
  return s;
}

}  // namespace dj_bootstrap
