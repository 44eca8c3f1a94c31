// fw25_cli.cu -- `fw25_engine`: executable drop-in for the reference's pre-compiled engine binaries.
//
// Honours the reference's file protocol exactly (the files
// /root/reference/fullwave/solver/input_file_writer.py:563-881 writes): run with no arguments in a directory
// of .dat files, read them, step, write genout.dat (float32 [n_frames][ncoordsout]), exit 0 / non-zero.  So
// `fullwave.Solver(..., path_fullwave_simulation_bin=<this file>)` works with zero edits upstream -- the
// reference copies the executable into the simulation directory and runs it there (input_file_writer.py:
// 823-827, launcher.py:196-215), which is why this binary links the engine statically instead of libfw25.so.
// CUDA_VISIBLE_DEVICES selects the GPUs like it does for the reference binary (launcher.py:206): every visible
// device takes one x-slab (the reference's `cuda_device_id=[0, 1, ...]`), as long as slabs stay >= 16 planes.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/fw25.h"

namespace {

bool file_exists(const char *name) {
  FILE *f = fopen(name, "rb");
  if (!f) return false;
  fclose(f);
  return true;
}

template <class T>
bool read_all(const std::string &name, std::vector<T> &out, size_t count, bool required = true) {
  out.assign(count, T(0));
  if (count == 0) return true;
  FILE *f = fopen(name.c_str(), "rb");
  if (!f) {
    if (required) fprintf(stderr, "fw25_engine: cannot open %s\n", name.c_str());
    return !required;
  }
  const size_t got = fread(out.data(), sizeof(T), count, f);
  fclose(f);
  if (got != count) {
    fprintf(stderr, "fw25_engine: %s holds %zu values, expected %zu\n", name.c_str(), got, count);
    return false;
  }
  printf("Reading data from file %s (size: %zu bytes)...\n", name.c_str(), count * sizeof(T));
  return true;
}

// Big arrays are mapped, not copied: the engine uploads straight from the page cache (the reference's loaders
// fread every file into a malloc'ed copy first; SURVEY.md 2.1 import_float_data).
struct Mapped {
  const void *ptr = nullptr;
  size_t bytes = 0;
  ~Mapped() { if (ptr && bytes) munmap(const_cast<void *>(ptr), bytes); }
};

template <class T>
const T *map_array(const std::string &name, size_t count, std::vector<Mapped *> &keep, bool required = true) {
  static const T dummy = T(0);
  if (count == 0) return &dummy;
  const int fd = open(name.c_str(), O_RDONLY);
  if (fd < 0) {
    if (required) fprintf(stderr, "fw25_engine: cannot open %s\n", name.c_str());
    return nullptr;
  }
  struct stat st;
  if (fstat(fd, &st) != 0 || (size_t)st.st_size < count * sizeof(T)) {
    fprintf(stderr, "fw25_engine: %s holds %zu values, expected %zu\n", name.c_str(),
            (size_t)st.st_size / sizeof(T), count);
    close(fd);
    return nullptr;
  }
  void *m = mmap(nullptr, count * sizeof(T), PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
  close(fd);
  if (m == MAP_FAILED) {
    fprintf(stderr, "fw25_engine: cannot map %s\n", name.c_str());
    return nullptr;
  }
  auto *h = new Mapped();
  h->ptr = m;
  h->bytes = count * sizeof(T);
  keep.push_back(h);
  printf("Reading data from file %s (size: %zu bytes)...\n", name.c_str(), count * sizeof(T));
  return static_cast<const T *>(m);
}

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

bool read_i32(const char *stem, int32_t &v, bool required = true) {
  std::vector<int32_t> t;
  if (!file_exists((std::string(stem) + ".dat").c_str()) && !required) { v = 0; return true; }
  if (!read_all(std::string(stem) + ".dat", t, 1)) return false;
  v = t[0];
  return true;
}

bool read_f32(const char *stem, float &v) {
  std::vector<float> t;
  if (!read_all(std::string(stem) + ".dat", t, 1)) return false;
  v = t[0];
  return true;
}

}  // namespace

int main() {
  setvbuf(stdout, nullptr, _IOLBF, 0);
  fw25_problem pb;
  memset(&pb, 0, sizeof pb);
  pb.ndim = file_exists("nZ.dat") ? 3 : 2;
  pb.nZ = 1;
  bool ok = read_i32("nX", pb.nX) && read_i32("nY", pb.nY) && (pb.ndim == 2 || read_i32("nZ", pb.nZ)) &&
            read_i32("nT", pb.nT) && read_i32("nTic", pb.nTic) && read_i32("modT", pb.modT) &&
            read_f32("dX", pb.dX) && read_f32("dT", pb.dT) && read_i32("ncoords", pb.ncoords) &&
            read_i32("ncoordszero", pb.ncoordszero, false) && read_i32("ncoordsout", pb.ncoordsout) &&
            read_i32("ndmap", pb.ndmap);
  if (!ok) return 2;
  printf("fw25_engine: %dD nX=%d nY=%d nZ=%d nT=%d nTic=%d modT=%d ncoords=%d ncoordszero=%d ncoordsout=%d ndmap=%d\n",
         pb.ndim, pb.nX, pb.nY, pb.nZ, pb.nT, pb.nTic, pb.modT, pb.ncoords, pb.ncoordszero, pb.ncoordsout, pb.ndmap);
  if (pb.nX <= 0 || pb.nY <= 0 || pb.nZ <= 0 || pb.ndmap <= 0 || pb.ncoords < 0 || pb.ncoordsout < 0 ||
      pb.ncoordszero < 0 || pb.nTic < 0) {
    fprintf(stderr, "fw25_engine: invalid scalar inputs\n");
    return 2;
  }
  const double t_start = now_ms();
  const size_t n = (size_t)pb.nX * pb.nY * pb.nZ;
  const char *names[13] = {"rho", "K", "beta", "kappax", "kappau", "apmlx1", "bpmlx1", "apmlx2", "bpmlx2",
                           "apmlu1", "bpmlu1", "apmlu2", "bpmlu2"};
  std::vector<Mapped *> keep;
  const float *maps[13];
  for (int i = 0; i < 13; ++i)
    if (!(maps[i] = map_array<float>(std::string(names[i]) + ".dat", n, keep))) return 2;
  // kappay.dat present: the anisotropic file set (input_file_writer.py:592-620) -- per-axis maps
  const bool aniso = file_exists("kappay.dat");
  fw25_aniso an;
  memset(&an, 0, sizeof an);
  if (aniso) {
    const char *vel = "xyz";
    const char *prs = pb.ndim == 2 ? "uw" : "uvw";
    auto load = [&](const std::string &stem) { return map_array<float>(stem + ".dat", n, keep); };
    bool ok2 = true;
    for (int ax = 0; ax < pb.ndim && ok2; ++ax) {
      const std::string v(1, vel[ax]), q(1, prs[ax]);
      ok2 = (an.kappa_vel[ax] = load("kappa" + v)) && (an.kappa_prs[ax] = load("kappa" + q));
      for (int nu = 0; nu < 2 && ok2; ++nu) {
        const std::string k = std::to_string(nu + 1);
        ok2 = (an.a_vel[ax][nu] = load("apml" + v + k)) && (an.b_vel[ax][nu] = load("bpml" + v + k)) &&
              (an.a_prs[ax][nu] = load("apml" + q + k)) && (an.b_prs[ax][nu] = load("bpml" + q + k));
      }
    }
    if (!ok2) return 2;
    pb.aniso = &an;
  }
  pb.dmap = map_array<float>("dmap.dat", (size_t)18 * pb.ndmap, keep);
  pb.dcmap = map_array<int32_t>("dcmap.dat", n, keep);
  pb.icc = map_array<int32_t>("icc.dat", (size_t)pb.ncoords * pb.ndim, keep);
  pb.icmat = map_array<float>("icmat.dat", (size_t)pb.ncoords * pb.nTic, keep);
  pb.outc = map_array<int32_t>("outc.dat", (size_t)pb.ncoordsout * pb.ndim, keep);
  pb.icczero = map_array<int32_t>("icczero.dat", (size_t)pb.ncoordszero * pb.ndim, keep, pb.ncoordszero > 0);
  if (!pb.dmap || !pb.dcmap || !pb.icc || !pb.icmat || !pb.outc || !pb.icczero) return 2;
  pb.rho = maps[0]; pb.K = maps[1]; pb.beta = maps[2];
  pb.kappax = maps[3]; pb.kappau = maps[4];
  pb.apmlx1 = maps[5]; pb.bpmlx1 = maps[6]; pb.apmlx2 = maps[7]; pb.bpmlx2 = maps[8];
  pb.apmlu1 = maps[9]; pb.bpmlu1 = maps[10]; pb.apmlu2 = maps[11]; pb.bpmlu2 = maps[12];

  // genout.dat is written through a shared mapping: the frames land in the page cache once, straight from the
  // device-to-host copies (no zero-filled staging vector, no fwrite pass)
  const size_t n_frames = pb.nT > 0 ? ((size_t)pb.nT + pb.modT - 1) / pb.modT : 0;
  const size_t n_out = n_frames * (size_t)pb.ncoordsout;
  float *genout = nullptr;
  std::vector<float> genout_fallback;
  {
    const int fd = open("genout.dat", O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) { fprintf(stderr, "fw25_engine: cannot create genout.dat\n"); return 3; }
    if (n_out > 0 && ftruncate(fd, (off_t)(n_out * sizeof(float))) == 0) {
      void *m = mmap(nullptr, n_out * sizeof(float), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
      if (m != MAP_FAILED) genout = static_cast<float *>(m);
    }
    close(fd);
    if (n_out > 0 && !genout) { genout_fallback.resize(n_out); }
  }
  float *out_ptr = genout ? genout : genout_fallback.data();
  const double t_read = now_ms();
  fw25_stats st;
  memset(&st, 0, sizeof st);
  int n_dev = fw25_device_count();
  if (n_dev < 1) n_dev = 1;                       // fw25_run reports the CUDA error
  while (n_dev > 1 && pb.nX / n_dev < 2 * FW25_M) --n_dev;
  std::vector<int32_t> devs(n_dev);
  for (int i = 0; i < n_dev; ++i) devs[i] = i;
  // FW25_DEVICE_LIST="0,0": explicit slab -> device assignment (tests: several slabs on ONE device, which
  // CUDA_VISIBLE_DEVICES cannot express); slabs thinner than two halos are still refused by fw25_run
  if (const char *dl = getenv("FW25_DEVICE_LIST")) {
    devs.clear();
    for (const char *q = dl; *q;) {
      char *end = nullptr;
      const long v = strtol(q, &end, 10);
      if (end == q) break;
      devs.push_back((int32_t)v);
      q = *end == ',' ? end + 1 : end;
    }
    if (devs.empty()) devs.push_back(0);
    n_dev = (int)devs.size();
  }
  printf("fw25_engine: %d GPU(s)\n", n_dev);
  const int rc = fw25_run(&pb, devs.data(), n_dev, out_ptr, n_out, &st);
  if (rc != 0) {
    fprintf(stderr, "fw25_engine: error %d: %s\n", rc, fw25_last_error());
    if (genout) munmap(genout, n_out * sizeof(float));
    unlink("genout.dat");
    return rc;
  }
  const double t_run = now_ms();
  if (genout) {
    munmap(genout, n_out * sizeof(float));
  } else if (n_out > 0) {
    FILE *f = fopen("genout.dat", "wb");
    if (!f || fwrite(genout_fallback.data(), sizeof(float), n_out, f) != n_out) {
      fprintf(stderr, "fw25_engine: cannot write genout.dat\n");
      return 3;
    }
    fclose(f);
  }
  for (auto *m : keep) delete m;
  printf("fw25_engine: inputs mapped in %.1f ms, engine %.1f ms (setup %.1f, loop %.1f, frames to host %.1f), output "
         "closed in %.1f ms\n", t_read - t_start, t_run - t_read, st.setup_ms, st.loop_ms, st.d2h_ms, now_ms() - t_run);
  printf("Progress : 1.000\nfw25_engine: %lld point-updates in %.3f ms (%.2f Gpt/s), setup %.1f ms, %lld kernel launches\n",
         (long long)st.point_updates, st.loop_ms, st.loop_ms > 0 ? st.point_updates / st.loop_ms / 1e6 : 0.0, st.setup_ms,
         (long long)st.kernel_launches);
  return 0;
}
