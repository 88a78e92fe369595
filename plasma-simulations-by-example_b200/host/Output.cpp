// Output.cpp -- VTK / screen / CSV output in the reference's formats (ch3/ver2/Output.cpp:12-133), reading the lazily
// synchronised host mirrors: a field is copied from the GPU only when a file that needs it is written.
#include "Output.h"

#include <iomanip>
#include <iostream>
#include <sstream>

namespace {
template <typename F>
void data_array(std::ostream &out, const std::string &name, int ncomp, const char *type, F &field)
{
    out << "<DataArray Name=\"" << name << "\" NumberOfComponents=\"" << ncomp << "\" format=\"ascii\" type=\"" << type << "\">\n";
    out << field;
    out << "</DataArray>\n";
}
std::ofstream f_diag;
}  // namespace

void Output::fields(World &world, std::vector<Species> &species)
{
    // ch4 (ch4/Output.cpp:12-118): a program that samples velocity moments also gets stream velocity, temperature and the
    // macroparticles-per-cell array, and its samples are restarted after every file until steady state
    bool moments = false;
    for (Species &sp : species) moments = moments || sp.samplesMoments();
    if (moments)
        for (Species &sp : species) sp.computeGasProperties();
    std::stringstream name;
    name << "results/fields_" << std::setfill('0') << std::setw(5) << world.getTs() << ".vti";
    std::ofstream out(name.str());
    if (!out.is_open()) { std::cerr << "Could not open " << name.str() << std::endl; return; }

    double3 x0 = world.getX0(), dh = world.getDh();
    out << "<VTKFile type=\"ImageData\">\n";
    out << "<ImageData Origin=\"" << x0[0] << " " << x0[1] << " " << x0[2] << "\" ";
    out << "Spacing=\"" << dh[0] << " " << dh[1] << " " << dh[2] << "\" ";
    out << "WholeExtent=\"0 " << world.ni - 1 << " 0 " << world.nj - 1 << " 0 " << world.nk - 1 << "\">\n";
    out << "<PointData>\n";
    data_array(out, "object_id", 1, "Int32", world.object_id);
    data_array(out, "NodeVol", 1, "Float64", world.node_vol);
    data_array(out, "phi", 1, "Float64", world.phi);
    data_array(out, "rho", 1, "Float64", world.rho);
    for (Species &sp : species) data_array(out, "nd." + sp.name, 1, "Float64", sp.den);
    for (Species &sp : species) data_array(out, "nd-ave." + sp.name, 1, "Float64", sp.den_ave);
    if (moments) {
        for (Species &sp : species) data_array(out, "vel." + sp.name, 3, "Float64", sp.vel);
        for (Species &sp : species) data_array(out, "T." + sp.name, 1, "Float64", sp.T);
    }
    data_array(out, "ef", 3, "Float64", world.ef);
    out << "</PointData>\n";
    if (moments) {
        out << "<CellData>\n";
        for (Species &sp : species) data_array(out, "mpc." + sp.name, 1, "Float64", sp.mpc);
        out << "</CellData>\n";
    }
    out << "</ImageData>\n</VTKFile>\n";
    out.close();
    if (moments && !world.isSteadyState())
        for (Species &sp : species) sp.clearSamples();
}

void Output::screenOutput(World &world, std::vector<Species> &species)
{
    std::cout << "ts: " << world.getTs();
    for (Species &sp : species) std::cout << std::setprecision(3) << "\t " << sp.name << ":" << sp.getNp();
    std::cout << std::endl;
}

void Output::diagOutput(World &world, std::vector<Species> &species)
{
    if (!f_diag.is_open()) {
        f_diag.open("runtime_diags.csv");
        f_diag << "ts,time,wall_time";
        for (Species &sp : species)
            for (const char *col : {"mp_count", "real_count", "px", "py", "pz", "KE"}) f_diag << "," << col << "." << sp.name;
        f_diag << ",PE,E_total" << std::endl;
    }
    f_diag << world.getTs() << "," << world.getTime() << "," << world.getWallTime();
    double tot_KE = 0;
    for (Species &sp : species) {
        double KE = sp.getKE();
        double3 mom = sp.getMomentum();
        tot_KE += KE;
        f_diag << "," << sp.getNp() << "," << sp.getRealCount() << "," << mom[0] << "," << mom[1] << "," << mom[2] << "," << KE;
    }
    double PE = world.getPE();
    f_diag << "," << PE << "," << (tot_KE + PE) << "\n";
    if (world.getTs() % 25 == 0) f_diag.flush();
}

// ch4/Output.cpp:175-229.  The selection rule is the reference's: a counter grows by num_parts/np per particle, a particle is
// written when it exceeds 1 and the counter restarts at -1 (so about num_parts/2 particles are written).
void Output::particlesVTP(std::ostream &out, const std::string &species_name, std::vector<Particle> &parts, int num_parts)
{
    const double dp = num_parts / (double)parts.size();
    double counter = 0;
    std::vector<Particle *> to_output;
    for (Particle &part : parts) {
        counter += dp;
        if (counter > 1) { to_output.emplace_back(&part); counter = -1; }
    }
    out << "<?xml version=\"1.0\"?>\n";
    out << "<VTKFile type=\"PolyData\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
    out << "<PolyData>\n";
    out << "<Piece NumberOfPoints=\"" << to_output.size() << "\" NumberOfVerts=\"0\" NumberOfLines=\"0\" ";
    out << "NumberOfStrips=\"0\" NumberOfCells=\"0\">\n";
    out << "<Points>\n";
    out << "<DataArray type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
    for (Particle *part : to_output) out << part->pos << "\n";
    out << "</DataArray>\n";
    out << "</Points>\n";
    out << "<PointData>\n";
    out << "<DataArray Name=\"vel." << species_name << "\" type=\"Float64\" NumberOfComponents=\"3\" format=\"ascii\">\n";
    for (Particle *part : to_output) out << part->vel << "\n";
    out << "</DataArray>\n";
    out << "</PointData>\n";
    out << "</Piece>\n";
    out << "</PolyData>\n";
    out << "</VTKFile>\n";
}

void Output::particles(World &world, std::vector<Species> &species, int num_parts)
{
    for (Species &sp : species) {
        std::stringstream name;
        name << "results/parts_" << sp.name << "_" << std::setfill('0') << std::setw(5) << world.getTs() << ".vtp";
        std::ofstream out(name.str());
        if (!out.is_open()) { std::cerr << "Could not open " << name.str() << std::endl; return; }
        std::vector<Particle> snapshot = sp.downloadParticles();     // particles live on the GPU: one device-to-host copy per file
        particlesVTP(out, sp.name, snapshot, num_parts);
        out.close();
    }
}
