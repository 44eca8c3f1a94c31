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
  const size_t n = (size_t)pb.nX * pb.nY * pb.nZ;
  const char *names[13] = {"rho", "K", "beta", "kappax", "kappau", "apmlx1", "bpmlx1", "apmlx2", "bpmlx2",
                           "apmlu1", "bpmlu1", "apmlu2", "bpmlu2"};
  // kappay.dat present: the anisotropic file set (input_file_writer.py:592-620) -- per-axis maps
  const bool aniso = file_exists("kappay.dat");
  std::vector<float> maps[13];
  for (int i = 0; i < 13; ++i)
    if (!read_all(std::string(names[i]) + ".dat", maps[i], n)) return 2;
  std::vector<std::vector<float>> amaps;
  fw25_aniso an;
  memset(&an, 0, sizeof an);
  if (aniso) {
    const char *vel = "xyz";
    const char *prs = pb.ndim == 2 ? "uw" : "uvw";
    amaps.reserve(60);
    auto load = [&](const std::string &stem) -> const float * {
      amaps.emplace_back();
      if (!read_all(stem + ".dat", amaps.back(), n)) return nullptr;
      return amaps.back().data();
    };
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
  std::vector<float> dmap, icmat;
  std::vector<int32_t> dcmap, icc, outc, icczero;
  if (!read_all("dmap.dat", dmap, (size_t)18 * pb.ndmap) || !read_all("dcmap.dat", dcmap, n) ||
      !read_all("icc.dat", icc, (size_t)pb.ncoords * pb.ndim) ||
      !read_all("icmat.dat", icmat, (size_t)pb.ncoords * pb.nTic) ||
      !read_all("outc.dat", outc, (size_t)pb.ncoordsout * pb.ndim) ||
      !read_all("icczero.dat", icczero, (size_t)pb.ncoordszero * pb.ndim, pb.ncoordszero > 0))
    return 2;
  pb.rho = maps[0].data(); pb.K = maps[1].data(); pb.beta = maps[2].data();
  pb.kappax = maps[3].data(); pb.kappau = maps[4].data();
  pb.apmlx1 = maps[5].data(); pb.bpmlx1 = maps[6].data(); pb.apmlx2 = maps[7].data(); pb.bpmlx2 = maps[8].data();
  pb.apmlu1 = maps[9].data(); pb.bpmlu1 = maps[10].data(); pb.apmlu2 = maps[11].data(); pb.bpmlu2 = maps[12].data();
  pb.dmap = dmap.data(); pb.dcmap = dcmap.data();
  pb.icc = icc.data(); pb.icmat = icmat.data(); pb.outc = outc.data(); pb.icczero = icczero.data();

  const size_t n_frames = pb.nT > 0 ? ((size_t)pb.nT + pb.modT - 1) / pb.modT : 0;
  std::vector<float> genout(n_frames * (size_t)pb.ncoordsout);
  fw25_stats st;
  int n_dev = fw25_device_count();
  if (n_dev < 1) n_dev = 1;                       // fw25_run reports the CUDA error
  while (n_dev > 1 && pb.nX / n_dev < 2 * FW25_M) --n_dev;
  std::vector<int32_t> devs(n_dev);
  for (int i = 0; i < n_dev; ++i) devs[i] = i;
  printf("fw25_engine: %d GPU(s)\n", n_dev);
  const int rc = fw25_run(&pb, devs.data(), n_dev, genout.data(), genout.size(), &st);
  if (rc != 0) {
    fprintf(stderr, "fw25_engine: error %d: %s\n", rc, fw25_last_error());
    return rc;
  }
  FILE *f = fopen("genout.dat", "wb");
  if (!f || fwrite(genout.data(), sizeof(float), genout.size(), f) != genout.size()) {
    fprintf(stderr, "fw25_engine: cannot write genout.dat\n");
    return 3;
  }
  fclose(f);
  printf("Progress : 1.000\nfw25_engine: %lld point-updates in %.3f ms (%.2f Gpt/s), setup %.1f ms, %lld kernel launches\n",
         (long long)st.point_updates, st.loop_ms, st.loop_ms > 0 ? st.point_updates / st.loop_ms / 1e6 : 0.0, st.setup_ms,
         (long long)st.kernel_launches);
  return 0;
}
