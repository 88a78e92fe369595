// World.cpp -- host side of the domain: owns the espic_ctx and keeps the public Field mirrors coherent with it.
// Reference behaviour followed: ch3/ver2/World.cpp:14-157 (cited per function).
#include "World.h"

#include <cmath>
#include <cstdlib>
#include <iostream>

#include "Species.h"

Rnd rnd;

Rnd::Rnd() : rnd_dist{0, 1.0}
{
    const char *s = std::getenv("ESPIC_SEED");
    seed_ = s ? std::strtoull(s, nullptr, 10) : ((unsigned long long)std::random_device()() << 32 | std::random_device()());
    mt_gen.seed((std::mt19937::result_type)(seed_ ^ (seed_ >> 32)));
}

// World::World (World.cpp:14-19).  node_vol is sized (ni,nj,nk): the reference's (ni,nk,nk) is a typo that only
// works because nk >= nj in its cases.
World::World(int ni, int nj, int nk)
    : ni{ni}, nj{nj}, nk{nk}, nn{ni, nj, nk}, phi(ni, nj, nk), rho(ni, nj, nk), node_vol(ni, nj, nk), ef(ni, nj, nk),
      object_id(ni, nj, nk)
{
    time_start = std::chrono::high_resolution_clock::now();
}

World::~World()
{
    if (ctx) espic_destroy(ctx);
}

// World::setExtents (World.cpp:22-36): the engine computes dh, the centroid and the node volumes on the device
void World::setExtents(const double3 _x0, const double3 _xm)
{
    if (ctx) {
        if (n_species) throw std::runtime_error("World::setExtents: cannot change the extents once species exist");
        espic_destroy(ctx);
        ctx = nullptr;
    }
    x0 = _x0;
    xm = _xm;
    const char *dev = std::getenv("ESPIC_DEVICE");
    espic_host::check(espic_create(&ctx, ni, nj, nk, x0.data(), xm.data(), dev ? std::atoi(dev) : 0), "espic_create");
    double h[3], c[3];
    espic_get_mesh(ctx, h, c);
    dh = double3(h);
    xc = double3(c);
    // anything the host wrote before (phi, object_id) wins; node volumes come from the device
    phi.bind(ctx, ESPIC_PHI, 0, false);
    rho.bind(ctx, ESPIC_RHO, 0, false);
    ef.bind(ctx, ESPIC_EF, 0, false);
    object_id.bind(ctx, ESPIC_OBJECT_ID, 0, false);
    node_vol.bind(ctx, ESPIC_NODE_VOL, 0, true);
}

espic_ctx *World::engine()
{
    if (!ctx) throw std::runtime_error("World: setExtents() must be called before the domain is used");
    return ctx;
}

void World::fields_to_device()
{
    phi.to_device();
    rho.to_device();
    ef.to_device();
    object_id.to_device();
}

int World::register_species(Species *, double mass, double charge, double mpw0)
{
    int id = espic_species_create(engine(), mass, charge, mpw0, 0);
    espic_host::check(id, "espic_species_create");
    n_species++;
    if (charge != 0) n_charged++;
    return id;
}

double World::getWallTime()
{
    std::chrono::duration<double> d = std::chrono::high_resolution_clock::now() - time_start;
    return d.count();
}

// World::computeChargeDensity (World.cpp:46-54).  The engine sums charge*den over every species of this World; the reference sums
// over the vector it is given and skips neutral species, so the two agree exactly when the vector holds every CHARGED species.
void World::computeChargeDensity(std::vector<Species> &species)
{
    int charged = 0;
    for (Species &sp : species) charged += sp.charge != 0;
    if (charged != n_charged)
        throw std::runtime_error("World::computeChargeDensity: pass every charged species created on this World");
    for (Species &sp : species) sp.den.to_device();
    espic_host::check(espic_charge_density(engine()), "espic_charge_density");
    rho.mark_device_wrote();
}

// World::getPE (World.cpp:72-84)
double World::getPE()
{
    ef.to_device();
    double pe = 0;
    espic_host::check(espic_field_pe(engine(), &pe), "espic_field_pe");
    return pe;
}

// World::addSphere (World.cpp:87-105)
void World::addSphere(double3 c, double radius, double phi_sphere)
{
    sphere_x0 = c;
    sphere_rad2 = radius * radius;
    fields_to_device();
    espic_host::check(espic_add_sphere(engine(), c.data(), radius, phi_sphere), "espic_add_sphere");
    object_id.mark_device_wrote();
    phi.mark_device_wrote();
}

// World::addInlet (World.cpp:108-115)
void World::addInlet()
{
    fields_to_device();
    espic_host::check(espic_add_inlet(engine()), "espic_add_inlet");
    object_id.mark_device_wrote();
    phi.mark_device_wrote();
}

// World::inSphere (World.cpp:118-125)
bool World::inSphere(double3 x)
{
    double3 r = x - sphere_x0;
    double r_mag2 = (r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
    return r_mag2 <= sphere_rad2;
}

// World::steadyState (World.cpp:128-157): relative change of mass, |p_z| and total energy below 1e-3
bool World::steadyState(std::vector<Species> &species)
{
    if (steady_state) return true;
    double tot_mass = 0, tot_mom = 0, tot_en = getPE();
    for (Species &sp : species) {
        tot_mass += sp.getRealCount();
        double3 mom = sp.getMomentum();
        tot_mom += std::abs(mom[2]);
        tot_en += sp.getKE();
    }
    const double tol = 1e-3;
    if (std::abs((tot_mass - last_mass) / tot_mass) < tol && std::abs((tot_mom - last_mom) / tot_mom) < tol &&
        std::abs((tot_en - last_en) / tot_en) < tol) {
        steady_state = true;
        std::cout << "Steady state reached at time step " << ts << std::endl;
    }
    last_mass = tot_mass;
    last_mom = tot_mom;
    last_en = tot_en;
    return steady_state;
}
