// shim_check.cpp -- drives the engine ONLY through the reference-style class API (World / Species / ColdBeamSource /
// PotentialSolver / Output), the way the book's Main.cpp does, and dumps the final state in the tests' ESPICST1 format
// (tests/statefile.py) so that pytest can compare it with the CPU oracle.  Scenarios:
//   shim_check sphere <ni> <nj> <nk> <steps> <GS|PCG|QN> <tol> <out.state>     ch3/ver2/Main.cpp:16-91 flow
//   shim_check box <n> <ions_grid> <eles_grid> <steps> <out.state>             ch2/Main.cpp:14-75 flow
//   shim_check surface <steps> <QN|PCG> <out.state>                               ch4/Main.cpp:19-110 flow (ions + neutrals)
//   shim_check fieldio <out.state>                                             host writes to Field mirrors reach the GPU
//   shim_check vtp <in.bin> <num_parts> <out.vtp>                              Output::particlesVTP on a particle snapshot from a file
//                                                                              (int64 np; double part[7][np]); host only, no GPU needed
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include "Output.h"
#include "PotentialSolver.h"
#include "Source.h"
#include "Species.h"
#include "World.h"

using namespace Const;

namespace {

struct Dump {
    FILE *f;
    explicit Dump(const std::string &path) : f(fopen(path.c_str(), "wb"))
    {
        if (!f) { perror(path.c_str()); exit(2); }
    }
    ~Dump() { fclose(f); }
    void i32(int v) { fwrite(&v, 4, 1, f); }
    void i64(long long v) { fwrite(&v, 8, 1, f); }
    void f64(double v) { fwrite(&v, 8, 1, f); }
    void vec(double3 v) { for (int c = 0; c < 3; c++) f64(v(c)); }
    void field(Field &a) { const std::vector<double> &h = a.host(); fwrite(h.data(), 8, h.size(), f); }
};

void dump_state(const std::string &path, World &world, std::vector<Species> &species, int flags, double3 sphere_c,
                double sphere_r, double sphere_phi, double phi0, double Te0, double n0, bool converged)
{
    Dump d(path);
    fwrite("ESPICST1", 1, 8, d.f);
    d.i32(world.ni); d.i32(world.nj); d.i32(world.nk); d.i32(flags); d.i32((int)species.size()); d.i32(0);
    d.vec(world.getX0()); d.vec(world.getXm()); d.f64(world.getDt());
    d.vec(sphere_c); d.f64(sphere_r); d.f64(sphere_phi);
    d.f64(phi0); d.f64(Te0); d.f64(n0);
    d.field(world.phi); d.field(world.rho);
    { const std::vector<double3> &h = world.ef.host(); fwrite(h.data(), 24, h.size(), d.f); }
    d.field(world.node_vol);
    { const std::vector<int> &h = world.object_id.host(); fwrite(h.data(), 4, h.size(), d.f); }
    double diag[16] = {0};
    diag[0] = converged ? 1.0 : 0.0;
    diag[1] = world.getPE();
    int q = 0;
    for (Species &sp : species) {
        std::vector<Particle> p = sp.downloadParticles();
        d.f64(sp.mass); d.f64(sp.charge); d.f64(sp.mpw0); d.i64((long long)p.size());
        d.field(sp.den); d.field(sp.den_ave);
        for (int c = 0; c < 7; c++)
            for (const Particle &pt : p) d.f64(c < 3 ? pt.pos(c) : (c < 6 ? pt.vel(c - 3) : pt.mpw));
        if (q < 2) {
            double3 mom = sp.getMomentum();
            diag[2 + 5 * q] = sp.getRealCount();
            for (int c = 0; c < 3; c++) diag[3 + 5 * q + c] = mom[c];
            diag[6 + 5 * q] = sp.getKE();
        }
        q++;
    }
    fwrite(diag, 8, 16, d.f);
}

int run_sphere(int argc, char **args)
{
    if (argc < 9) return 1;
    const int ni = atoi(args[2]), nj = atoi(args[3]), nk = atoi(args[4]), steps = atoi(args[5]);
    const std::string st = args[6];
    const double tol = atof(args[7]);
    World world(ni, nj, nk);
    world.setExtents({-0.1, -0.1, 0}, {0.1, 0.1, 0.4});
    world.setTime(1e-7, steps);
    const double3 sc{0, 0, 0.15};
    world.addSphere(sc, 0.05, -100);
    world.addInlet();

    std::vector<Species> species;
    species.push_back(Species("O+", 16 * AMU, QE, 2e2, world));
    const double ndi = 1e10;
    std::vector<ColdBeamSource> sources;
    sources.push_back(ColdBeamSource(species[0], world, 7000, ndi));

    PotentialSolver solver(world, st == "GS" ? SolverType::GS : (st == "PCG" ? SolverType::PCG : SolverType::QN), 20000, tol);
    solver.setReferenceValues(0, 1.5, ndi);
    bool ok = solver.solve();
    solver.computeEF();
    while (world.advanceTime()) {
        for (ColdBeamSource &source : sources) source.sample();
        for (Species &sp : species) {
            sp.advance();
            sp.computeNumberDensity();
        }
        world.computeChargeDensity(species);
        ok = solver.solve();
        solver.computeEF();
        if (world.steadyState(species) || world.getTs() >= steps / 2)
            for (Species &sp : species) sp.updateAverages();
        Output::screenOutput(world, species);
        Output::diagOutput(world, species);
        if (world.isLastTimeStep()) {
            Output::fields(world, species);
            // the reference's loop runs one more step after its "last" one (advanceTime: ts <= num_ts, isLastTimeStep: ts == num_ts-1,
            // World.h): keep the state the file was written from, for the comparison with the reference's own writer
            dump_state(std::string(args[8]) + ".fields", world, species, 3, sc, 0.05, -100, 0, 1.5, ndi, ok);
        }
    }
    dump_state(args[8], world, species, 3, sc, 0.05, -100, 0, 1.5, ndi, ok);
    return 0;
}

// ch4/Main.cpp flow with its commented-out ion species switched on: warm neutral beam + cold ion beam, every species advanced with
// surface interactions (neutrals bounce off the sphere, ions that hit it come back as neutrals), moments sampled every step
int run_surface(int argc, char **args)
{
    if (argc < 5) return 1;
    const int steps = atoi(args[2]);
    const std::string st = args[3];
    World world(21, 21, 41);
    world.setExtents({-0.1, -0.1, 0}, {0.1, 0.1, 0.4});
    world.setTime(2e-7, steps);
    const double3 sc{0, 0, 0.15};
    world.addSphere(sc, 0.05, -100);
    world.addInlet();
    std::vector<Species> species;
    species.reserve(2);
    species.push_back(Species("O", 16 * AMU, 0, 50, world));       // light neutral macroparticles: every ion impact emits two
    species.push_back(Species("O+", 16 * AMU, QE, 1e2, world));
    Species &neutrals = species[0];
    Species &ions = species[1];
    const double nda = 1e10, ndi = 1e10;
    std::vector<std::unique_ptr<Source>> sources;
    sources.emplace_back(new WarmBeamSource(neutrals, world, 7000, nda, 1000));
    sources.emplace_back(new ColdBeamSource(ions, world, 7000, ndi));
    PotentialSolver solver(world, st == "PCG" ? SolverType::PCG : SolverType::QN, 1000, 1e-4);
    solver.setReferenceValues(0, 1.5, ndi);
    bool ok = solver.solve();
    solver.computeEF();
    double emitted_weight = 0;
    while (world.advanceTime()) {
        for (auto &source : sources) source->sample();
        for (Species &sp : species) {
            const double before = neutrals.getRealCount();
            sp.advance(neutrals, neutrals);
            if (&sp == &ions) emitted_weight += neutrals.getRealCount() - before;
            sp.computeNumberDensity();
            sp.sampleMoments();
        }
        world.computeChargeDensity(species);
        ok = solver.solve();
        solver.computeEF();
        Output::screenOutput(world, species);
    }
    for (Species &sp : species) sp.computeGasProperties();
    printf("emitted neutral weight = %.15g\n", emitted_weight);
    dump_state(args[4], world, species, 3, sc, 0.05, -100, 0, 1.5, ndi, ok);
    return 0;
}

int run_box(int argc, char **args)
{
    if (argc < 7) return 1;
    const int n = atoi(args[2]), gi = atoi(args[3]), ge = atoi(args[4]), steps = atoi(args[5]);
    World world(n, n, n);
    world.setExtents({-0.1, -0.1, 0}, {0.1, 0.1, 0.2});
    world.setTime(2e-10, steps);
    std::vector<Species> species;
    species.reserve(2);
    species.push_back(Species("O+", 16 * AMU, QE, world));
    species.push_back(Species("e-", ME, -1 * QE, world));
    int3 np_ions_grid = {gi, gi, gi};
    int3 np_eles_grid = {ge, ge, ge};
    species[0].loadParticlesBoxQS(world.getX0(), world.getXm(), 1e11, np_ions_grid);
    species[1].loadParticlesBoxQS(world.getX0(), world.getXc(), 1e11, np_eles_grid);
    PotentialSolver solver(world, 10000, 1e-8);
    // the reference's ch2 Main solves before any density exists (rho = 0): phi stays 0
    bool ok = solver.solve();
    solver.computeEF();
    while (world.advanceTime()) {
        for (Species &sp : species) {
            sp.advance();
            sp.computeNumberDensity();
        }
        world.computeChargeDensity(species);
        ok = solver.solve();
        solver.computeEF();
        Output::screenOutput(world, species);
        Output::diagOutput(world, species);
    }
    dump_state(args[6], world, species, 0, double3(0, 0, 0), 0, 0, 0, 1.5, 1e12, ok);
    return 0;
}

// host-side writes through the Field mirrors must reach the device, and device results must come back
int run_fieldio(int argc, char **args)
{
    if (argc < 3) return 1;
    World world(7, 6, 9);
    world.setExtents({0, 0, 0}, {0.6, 0.5, 0.8});
    world.setTime(1e-9, 1);
    for (int i = 0; i < world.ni; i++)
        for (int j = 0; j < world.nj; j++)
            for (int k = 0; k < world.nk; k++) world.phi[i][j][k] = 3.0 * i * i - 2.0 * j + 0.5 * k * k * k + i * j * k;
    std::vector<Species> species;
    species.push_back(Species("O+", 16 * AMU, QE, 10.0, world));
    PotentialSolver solver(world, SolverType::QN, 1, 1.0);   // constructor's QN overwrites phi on free nodes with rho=0 -> log(1e-6)
    for (int i = 0; i < world.ni; i++)
        for (int j = 0; j < world.nj; j++)
            for (int k = 0; k < world.nk; k++) world.phi[i][j][k] = 3.0 * i * i - 2.0 * j + 0.5 * k * k * k + i * j * k;
    solver.computeEF();                       // device kernel must see the host-written phi
    double3 e = world.ef(3, 2, 4);            // and the host must see the device-written ef
    std::cout << "ef(3,2,4) = " << e << std::endl;
    species[0].addParticle({0.31, 0.22, 0.41}, {10, 20, 30}, 10.0);
    species[0].addParticle({0.7, 0.22, 0.41}, {10, 20, 30}, 10.0);     // out of bounds: dropped
    species[0].addParticle({0.05, 0.45, 0.79}, {-5, 0, 2}, 4.0);
    std::cout << "np = " << species[0].getNp() << std::endl;
    species[0].computeNumberDensity();
    world.computeChargeDensity(species);
    double sum = 0;
    for (int i = 0; i < world.ni; i++)
        for (int j = 0; j < world.nj; j++)
            for (int k = 0; k < world.nk; k++) sum += species[0].den(i, j, k) * world.node_vol(i, j, k);
    std::cout << "sum(den*vol) = " << sum << " real count = " << species[0].getRealCount() << std::endl;
    dump_state(args[2], world, species, 4, double3(0, 0, 0), 0, 0, 0, 1.5, 1e12, true);
    return 0;
}

}  // namespace

int run_vtp(int argc, char **args)
{
    if (argc < 5) return 1;
    FILE *f = fopen(args[2], "rb");
    if (!f) { perror(args[2]); return 2; }
    long long np = 0;
    if (fread(&np, 8, 1, f) != 1) return 2;
    std::vector<double> p((size_t)7 * np);
    if (fread(p.data(), 8, p.size(), f) != p.size()) return 2;
    fclose(f);
    std::vector<Particle> parts;
    for (long long i = 0; i < np; i++)
        parts.emplace_back(double3(p[i], p[np + i], p[2 * np + i]), double3(p[3 * np + i], p[4 * np + i], p[5 * np + i]), p[6 * np + i]);
    std::ofstream out(args[4]);
    Output::particlesVTP(out, "O+", parts, atoi(args[3]));
    return 0;
}

int main(int argc, char **args)
{
    int rc = 1;
    try {
        if (argc > 1 && !strcmp(args[1], "sphere")) rc = run_sphere(argc, args);
        else if (argc > 1 && !strcmp(args[1], "box")) rc = run_box(argc, args);
        else if (argc > 1 && !strcmp(args[1], "surface")) rc = run_surface(argc, args);
        else if (argc > 1 && !strcmp(args[1], "fieldio")) rc = run_fieldio(argc, args);
        else if (argc > 1 && !strcmp(args[1], "vtp")) rc = run_vtp(argc, args);
    } catch (const std::exception &e) {
        std::cerr << "shim_check: " << e.what() << std::endl;
        return 3;
    }
    if (rc == 1) std::cerr << "usage: shim_check sphere|box|surface|fieldio|vtp ... (see the header of shim_check.cpp)" << std::endl;
    return rc;
}
