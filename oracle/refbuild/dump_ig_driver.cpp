/* Golden-fixture dump driver for the initial-guess strategies (test infrastructure; build container only).
 * Links the UNMODIFIED reference and drives its InitialGuess classes (libs/linearSolver/initialGuess.cpp) the way
 * linearSolver_t::Solve does (FormInitialGuess -> solve -> Update) over a sequence of slowly varying right-hand sides
 * of a fixed SPD tridiagonal operator; the "solver" is an exact Thomas solve.  Dumps the right-hand sides, the
 * solutions and the initial guess every strategy forms at every step.
 * usage: dump_ig_driver outdir */
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>
#include "core.hpp"
#include "platform.hpp"
#include "initialGuess.hpp"

using namespace libp;

static const int N = 400, K = 14;

class tridiag_t : public operator_t {
 public:
  std::vector<double> d;
  tridiag_t() : d(N) { for (int i = 0; i < N; ++i) d[i] = 2.5 + 0.5 * std::sin(0.05 * i); }
  void Operator(deviceMemory<dfloat>& o_q, deviceMemory<dfloat>& o_Aq) {
    const dfloat* q = (const dfloat*)o_q.ptr();  // Serial mode: host pointers
    dfloat* Aq = (dfloat*)o_Aq.ptr();
    for (int i = 0; i < N; ++i)
      Aq[i] = d[i] * q[i] - (i > 0 ? q[i - 1] : 0.0) - (i + 1 < N ? q[i + 1] : 0.0);
  }
  void solve(const std::vector<double>& b, std::vector<double>& x) const {  // Thomas
    std::vector<double> c(N), g(N);
    c[0] = -1.0 / d[0]; g[0] = b[0] / d[0];
    for (int i = 1; i < N; ++i) {
      const double m = d[i] + c[i - 1];
      c[i] = -1.0 / m;
      g[i] = (b[i] + g[i - 1]) / m;
    }
    x.assign(N, 0.0);
    x[N - 1] = g[N - 1];
    for (int i = N - 2; i >= 0; --i) x[i] = g[i] - c[i] * x[i + 1];
  }
};

static void dumpv(const std::string& fn, const std::vector<double>& v) {
  FILE* f = fopen(fn.c_str(), "wb");
  fwrite(v.data(), sizeof(double), v.size(), f);
  fclose(f);
}

template <class IG>
static void run(const std::string& out, const std::string& name, platform_t& platform, settings_t& settings, comm_t comm,
                tridiag_t& A, const std::vector<std::vector<double>>& rhs, const std::vector<std::vector<double>>& sol) {
  IG ig(N, platform, settings, comm);
  std::vector<double> guesses;
  memory<dfloat> hx(N, 0.0), hb(N, 0.0);
  deviceMemory<dfloat> o_x = platform.malloc<dfloat>(hx);
  deviceMemory<dfloat> o_b = platform.malloc<dfloat>(hb);
  for (int k = 0; k < K; ++k) {
    for (int i = 0; i < N; ++i) hb[i] = rhs[k][i];
    o_b.copyFrom(hb);
    ig.FormInitialGuess(o_x, o_b);   // o_x enters holding the previous solution, as in a time stepper
    o_x.copyTo(hx);
    for (int i = 0; i < N; ++i) guesses.push_back(hx[i]);
    for (int i = 0; i < N; ++i) hx[i] = sol[k][i];
    o_x.copyFrom(hx);
    ig.Update(A, o_x, o_b);
  }
  dumpv(out + "/guess_" + name + ".f64.bin", guesses);
}

int main(int argc, char** argv) {
  Comm::Init(argc, argv);
  {
    comm_t comm(Comm::World().Dup());
    std::string out = argv[1];
    platformSettings_t ps(comm);
    ps.changeSetting("THREAD MODEL", "Serial");
    platform_t platform(ps);
    platform.linAlg().InitKernels({"set", "norm2", "axpy", "zaxpy", "innerProd"});
    tridiag_t A;
    std::vector<std::vector<double>> rhs(K, std::vector<double>(N)), sol(K);
    std::vector<double> flat_r, flat_x;
    for (int k = 0; k < K; ++k) {
      const double t = 0.1 * k;
      for (int i = 0; i < N; ++i) {
        const double s = (i + 0.5) / N;
        rhs[k][i] = std::sin(3.0 * s + t) + 0.3 * t * t * std::cos(7.0 * s) + 0.05 * std::sin(40.0 * s * t);
      }
      A.solve(rhs[k], sol[k]);
      flat_r.insert(flat_r.end(), rhs[k].begin(), rhs[k].end());
      flat_x.insert(flat_x.end(), sol[k].begin(), sol[k].end());
    }
    dumpv(out + "/rhs.f64.bin", flat_r);
    dumpv(out + "/sol.f64.bin", flat_x);
    dumpv(out + "/diag.f64.bin", A.d);
    auto mk = [&](int dim, int deg, const char* method) {
      settings_t s(comm);
      InitialGuess::AddSettings(s);
      s.changeSetting("INITIAL GUESS HISTORY SPACE DIMENSION", std::to_string(dim));
      s.changeSetting("INITIAL GUESS EXTRAP DEGREE", std::to_string(deg));
      s.changeSetting("INITIAL GUESS EXTRAP COEFFS METHOD", method);
      return s;
    };
    { settings_t s = mk(4, 2, "MINNORM"); run<InitialGuess::Zero>(out, "zero", platform, s, comm, A, rhs, sol); }
    { settings_t s = mk(4, 2, "MINNORM"); run<InitialGuess::ClassicProjection>(out, "classic4", platform, s, comm, A, rhs, sol); }
    { settings_t s = mk(5, 2, "MINNORM"); run<InitialGuess::RollingQRProjection>(out, "qr5", platform, s, comm, A, rhs, sol); }
    { settings_t s = mk(3, 2, "MINNORM"); run<InitialGuess::RollingQRProjection>(out, "qr3", platform, s, comm, A, rhs, sol); }
    { settings_t s = mk(4, 2, "MINNORM"); run<InitialGuess::Extrap>(out, "extrap_m2_M4_minnorm", platform, s, comm, A, rhs, sol); }
    { settings_t s = mk(6, 3, "MINNORM"); run<InitialGuess::Extrap>(out, "extrap_m3_M6_minnorm", platform, s, comm, A, rhs, sol); }
    { settings_t s = mk(5, 2, "CPQR"); run<InitialGuess::Extrap>(out, "extrap_m2_M5_cpqr", platform, s, comm, A, rhs, sol); }
  }
  Comm::Finalize();
  return 0;
}
