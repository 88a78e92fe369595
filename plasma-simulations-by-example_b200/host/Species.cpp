// Species.cpp -- forwards the reference's Species interface (ch3/ver2/Species.cpp, ch2/Species.cpp) to the C ABI.
#include "Species.h"

#include <cstdlib>

static int default_sort_every()
{
    const char *s = std::getenv("ESPIC_SORT_EVERY");
    return s ? std::atoi(s) : 0;
}

Species::Species(std::string name, double mass, double charge, double mpw0, World &world)
    : name(name), mass(mass), charge(charge), mpw0(mpw0), den(world.ni, world.nj, world.nk), den_ave(world.ni, world.nj, world.nk),
      T(world.ni, world.nj, world.nk), vel(world.ni, world.nj, world.nk), mpc(world.ni - 1, world.nj - 1, world.nk - 1), sort_every(default_sort_every()), world(world), reflect(false)
{
    sp_id = world.register_species(this, mass, charge, mpw0);
    bind_fields();
}

Species::Species(std::string name, double mass, double charge, World &world)
    : name(name), mass(mass), charge(charge), mpw0(0), den(world.ni, world.nj, world.nk), den_ave(world.ni, world.nj, world.nk),
      T(world.ni, world.nj, world.nk), vel(world.ni, world.nj, world.nk), mpc(world.ni - 1, world.nj - 1, world.nk - 1), sort_every(default_sort_every()), world(world), reflect(true)
{
    sp_id = world.register_species(this, mass, charge, 0);
    bind_fields();
}

Species::Species(Species &&o)
    : name(o.name), mass(o.mass), charge(o.charge), mpw0(o.mpw0), den(std::move(o.den)), den_ave(std::move(o.den_ave)),
      T(std::move(o.T)), vel(std::move(o.vel)), mpc(std::move(o.mpc)), sort_every(o.sort_every), world(o.world), sp_id(o.sp_id), reflect(o.reflect), n_advance(o.n_advance), diag_valid(o.diag_valid), moments_used(o.moments_used)
{
    for (int q = 0; q < 7; q++) pending[q] = std::move(o.pending[q]);
    for (int q = 0; q < 5; q++) diag[q] = o.diag[q];
    o.sp_id = -1;
}

void Species::bind_fields()
{
    den.bind(world.engine(), ESPIC_DEN, sp_id, true);
    den_ave.bind(world.engine(), ESPIC_DEN_AVE, sp_id, true);
}

// Species::addParticle (Species.cpp:65-81).  Queued on the host; the bounds test, the E gather and the half-step
// velocity rewind run on the device for the whole batch, in call order, when the particles are first needed.
void Species::addParticle(double3 pos, double3 vel, double mpwt)
{
    for (int c = 0; c < 3; c++) { pending[c].push_back(pos[c]); pending[3 + c].push_back(vel[c]); }
    pending[6].push_back(mpwt);
}

void Species::flush()
{
    if (pending[0].empty()) return;
    world.fields_to_device();            // the rewind gathers E
    const double *comp[7];
    for (int q = 0; q < 7; q++) comp[q] = pending[q].data();
    long long added = 0;
    espic_host::check(espic_species_add(world.engine(), sp_id, comp, (long long)pending[0].size(), world.getDt(), &added),
                      "espic_species_add");
    for (int q = 0; q < 7; q++) pending[q].clear();
    particles_changed();
}

size_t Species::getNp()
{
    flush();
    return (size_t)espic_species_count(world.engine(), sp_id);
}

void Species::refresh_diag()
{
    flush();
    if (diag_valid) return;
    espic_host::check(espic_species_diag(world.engine(), sp_id, diag), "espic_species_diag");
    diag_valid = true;
}

// Species.cpp:84-108: one fused device reduction serves all three
double Species::getRealCount() { refresh_diag(); return diag[0]; }
double3 Species::getMomentum() { refresh_diag(); return double3(diag[1], diag[2], diag[3]); }
double Species::getKE() { refresh_diag(); return diag[4]; }

// Species::advance (ch3/ver2/Species.cpp:7-48, ch2/Species.cpp:7-38)
void Species::advance()
{
    flush();
    world.fields_to_device();
    // with the periodic cell sort enabled the sort goes between push and scatter (the deposit then sees perfectly
    // ordered particles); otherwise the scatter is fused into the push kernel.  Absorbing walls (ch3, ch9): the push also
    // sums what getRealCount / getMomentum / getKE return (Species.cpp:84-108), so the diagnostics the book's Main.cpp prints
    // every step cost no second pass over the particles.
    const bool sort_now = sort_every > 0 && n_advance % sort_every == 0;
    n_advance++;
    int flags = sort_every > 0 ? 0 : ESPIC_PUSH_FUSE_DEPOSIT;
    if (!reflect) flags = ESPIC_PUSH_DIAG;
    espic_host::check(espic_push(world.engine(), sp_id, world.getDt(), reflect ? ESPIC_WALL_REFLECT : ESPIC_WALL_ABSORB, flags),
                      "espic_push");
    if (sort_now) sortByCell();
    particles_changed();
}

// ch4 Species::advance(neutrals, spherium) (ch4/Species.cpp:8-91).  Philox key: (world seed, 0x100 + species id, time step).
void Species::advance(Species &neutrals, Species &spherium)
{
    flush();
    neutrals.flush();
    spherium.flush();
    world.fields_to_device();
    n_advance++;
    long long emitted[2] = {0, 0};
    espic_host::check(espic_push_surface(world.engine(), sp_id, world.getDt(), neutrals.sp_id, spherium.sp_id, rnd.seed(),
                                         0x100u + (unsigned)sp_id, (unsigned)world.getTs(), emitted),
                      "espic_push_surface");
    particles_changed();
    if (emitted[0] || emitted[1]) { neutrals.particles_changed(); spherium.particles_changed(); }
}

// Species::computeNumberDensity (Species.cpp:51-62)
void Species::computeNumberDensity()
{
    flush();
    espic_host::check(espic_deposit(world.engine(), sp_id, ESPIC_DEPOSIT_FP64), "espic_deposit");
    den.mark_device_wrote();
}

void Species::updateAverages()
{
    den.to_device();
    den_ave.to_device();
    espic_host::check(espic_update_average(world.engine(), sp_id), "espic_update_average");
    den_ave.mark_device_wrote();
}

// ch4/Species.cpp:190-241, on the device
void Species::sampleMoments()
{
    flush();
    espic_host::check(espic_sample_moments(world.engine(), sp_id), "espic_sample_moments");
    moments_used = true;
}

void Species::computeGasProperties()
{
    espic_host::check(espic_compute_gas_properties(world.engine(), sp_id), "espic_compute_gas_properties");
    if (!T.bound()) {          // the moment arrays exist on the device from the first use on
        T.bind(world.engine(), ESPIC_T, sp_id, true);
        vel.bind(world.engine(), ESPIC_VEL, sp_id, true);
    }
    T.mark_device_wrote();
    vel.mark_device_wrote();
}

// ch4/Species.cpp:228-235
void Species::computeMPC()
{
    flush();
    espic_host::check(espic_compute_mpc(world.engine(), sp_id), "espic_compute_mpc");
    if (!mpc.bound()) mpc.bind(world.engine(), ESPIC_MPC, sp_id, true);
    mpc.mark_device_wrote();
}

void Species::clearSamples()
{
    espic_host::check(espic_clear_samples(world.engine(), sp_id), "espic_clear_samples");
}

void Species::sortByCell()
{
    flush();
    espic_host::check(espic_sort_by_cell(world.engine(), sp_id), "espic_sort_by_cell");
}

// ch2/Species.cpp:74-97
void Species::loadParticlesBox(double3 x1, double3 x2, double num_den, int num_mp)
{
    double box_vol = (x2[0] - x1[0]) * (x2[1] - x1[1]) * (x2[2] - x1[2]);
    double num_real = num_den * box_vol;
    double mpw = num_real / num_mp;
    for (int q = 0; q < 7; q++) pending[q].reserve(pending[q].size() + num_mp);
    for (int p = 0; p < num_mp; p++) {
        double3 pos;
        pos[0] = x1[0] + rnd() * (x2[0] - x1[0]);
        pos[1] = x1[1] + rnd() * (x2[1] - x1[1]);
        pos[2] = x1[2] + rnd() * (x2[2] - x1[2]);
        addParticle(pos, double3(0, 0, 0), mpw);
    }
}

// ch2/Species.cpp:101-141: (n-1)^3-weighted uniform grid, half weights on the faces, max-face particles nudged inside
void Species::loadParticlesBoxQS(double3 x1, double3 x2, double num_den, int3 num_mp)
{
    double box_vol = (x2[0] - x1[0]) * (x2[1] - x1[1]) * (x2[2] - x1[2]);
    int num_mp_tot = (num_mp[0] - 1) * (num_mp[1] - 1) * (num_mp[2] - 1);
    double num_real = num_den * box_vol;
    double mpw = num_real / num_mp_tot;
    double d[3];
    for (int a = 0; a < 3; a++) d[a] = (x2[a] - x1[a]) / (num_mp[a] - 1);
    const size_t total = (size_t)num_mp[0] * num_mp[1] * num_mp[2];
    for (int q = 0; q < 7; q++) pending[q].reserve(pending[q].size() + total);
    int n[3];
    for (n[0] = 0; n[0] < num_mp[0]; n[0]++)
        for (n[1] = 0; n[1] < num_mp[1]; n[1]++)
            for (n[2] = 0; n[2] < num_mp[2]; n[2]++) {
                double pos[3];
                double w = 1;
                for (int a = 0; a < 3; a++) {
                    pos[a] = x1[a] + n[a] * d[a];
                    if (pos[a] == x2[a]) pos[a] -= 1e-4 * d[a];
                    if (n[a] == 0 || n[a] == num_mp[a] - 1) w *= 0.5;
                }
                addParticle(double3(pos), double3(0, 0, 0), mpw * w);
            }
}

std::vector<Particle> Species::downloadParticles()
{
    flush();
    const long long n = espic_species_count(world.engine(), sp_id);
    std::vector<double> buf[7];
    double *comp[7];
    for (int q = 0; q < 7; q++) { buf[q].resize((size_t)n); comp[q] = buf[q].data(); }
    if (espic_species_download(world.engine(), sp_id, comp, n) < 0) espic_host::fail("espic_species_download");
    std::vector<Particle> out;
    out.reserve((size_t)n);
    for (long long i = 0; i < n; i++)
        out.emplace_back(double3(buf[0][i], buf[1][i], buf[2][i]), double3(buf[3][i], buf[4][i], buf[5][i]), buf[6][i]);
    return out;
}
