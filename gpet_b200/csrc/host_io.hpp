// Host-side input contract of gPET (file formats) -- parsed into plain structs.
// Mirrors the reference loaders: main.cu:50-184 (input_PET.in), detector.cu:64-285 (.geo),
// initialize.cu:10-144 (isotopes, phantom, psf, source), initialize.cu:279-748 (cross-section tables).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/gpet_b200.h"

namespace gpet {

// fgets/fscanf-compatible cursor over a whole file held in memory.
struct Scanner {
    std::string buf;
    size_t pos = 0;
    bool ok = true;

    bool load(const std::string& path);
    bool eof() const { return pos >= buf.size(); }
    void ws();                               // skip isspace() characters (what ' ' or '\n' in a scanf format do)
    std::string line(size_t maxlen = 0);     // fgets(buf, maxlen): at most maxlen-1 chars, through '\n'; 0 = unbounded
    bool i32(int32_t& v);                    // "%d" followed by whitespace skip
    bool u64(uint64_t& v);
    bool f32(float& v);                      // "%f" (strtof) followed by whitespace skip
    bool f64(double& v);
    bool word(std::string& v);               // "%s"
    void ignore_through(size_t n, char delim);  // istream::ignore(n, delim)
    bool next_is_number();                   // after ws(): does a numeric token start here?
    void skip_labels();                      // consume whole lines until the next token is numeric
};

struct Config {  // input_PET.in, 23 label/value groups (main.cu:52-182)
    int32_t device = 0;
    float nonangle = 0.f;
    int32_t pdim[3] = {0, 0, 0};
    float poffset[3] = {0, 0, 0}, psize[3] = {0, 0, 0};
    std::string matfile, denfile;
    int32_t nhist = 0, usepsf = 0;
    std::string sourcefile;
    int32_t ptype = -1, useprange = 0;
    float tstart = 0.f, tend = 1.f;
    float recordsphere[4] = {0, 0, 0, 0};
    float eabsph = 0.f;
    std::string geofile;
    int32_t nsurface = 0;
    std::vector<float> surface;
    int32_t rdepth = 0, rpolicy = 0;
    float Eth = 0.f;
    int32_t blurpolicy = 0;
    float Eref = 0.f, Rref = 0.f, Eslope = 0.f, Sblur = 0.f;
    int32_t dlevel = 0, dtype = 0;
    float dtime = 0.f;
    float Ewinmin = 0.f, Ewinmax = 0.f;
};

struct Geometry {  // read_file_ro + iniPanel-derived counts
    std::vector<gpet_panel> panels;
    int32_t mat[2] = {0, 0};   // crystal, gap
    float dens[2] = {0, 0};
    float rot_axis[3] = {0, 0, 1};
    float rot_angle_deg = 0.f;
    int32_t moduleNy = 0, crystalNy = 0, moduleN = 0, crystalN = 0;  // initialize.cu:1074-1086
};

struct Isotopes {  // data/isotopes.txt
    std::vector<float> halftime, ratio, coef;  // coef: 8 per isotope
    int n() const { return (int)halftime.size(); }
};

struct Sources {  // source.txt
    std::vector<uint64_t> natom;  // reference: unsigned int (gPET.h:50); widened
    std::vector<int32_t> type, shape;
    std::vector<float> coeff;     // 6 per source
    int n() const { return (int)natom.size(); }
};

struct Phantom {
    int32_t dim[3] = {0, 0, 0};
    float offset[3] = {0, 0, 0}, size[3] = {0, 0, 0}, d[3] = {0, 0, 0};
    std::vector<int32_t> mat;
    std::vector<float> dens;
    size_t nvox() const { return (size_t)dim[0] * dim[1] * dim[2]; }
};

struct Psf {  // readParticle
    std::vector<gpet_photon> p;  // eventid = record index, parn = record index
    int ptype = 1;
};

struct Tables {
    int32_t nmat = 0, nen = 0;
    float eminph = 0.f, emax = 0.f;            // .matter header
    std::vector<std::string> names;
    std::vector<float> refdens;
    std::vector<float> energy;                 // nen (shared grid of the four 1-D tables)
    std::vector<float> lamph, compt, phote, rayle;  // [mat][ie], cm^2/g
    int32_t cm_ncp = 0, cm_ne = 0, rl_ncp = 0, rl_ne = 0;
    float cm_dcp = 0.f, cm_de = 0.f, rl_dcp = 0.f, rl_de = 0.f;
    std::vector<float> cmpsf, rayff;           // [mat][icp][ie] cos(theta)
    bool loaded() const { return nmat > 0; }
};

// Each returns "" on success, else an error message.
std::string parse_config(const std::string& path, Config& out);
std::string parse_geometry(const std::string& path, Geometry& out);
std::string parse_isotopes(const std::string& path, Isotopes& out);
std::string parse_sources(const std::string& path, Sources& out);
std::string load_phantom(const std::string& matfile, const std::string& denfile, const int32_t dim[3],
                         const float offset[3], const float size[3], Phantom& out);
std::string load_psf(const std::string& path, int64_t max_particles, int ptype, Psf& out);
std::string load_tables_ascii(const std::string& prefix, Tables& out);
std::string load_tables_packed(const std::string& path, Tables& out);
std::string save_tables_packed(const std::string& path, const Tables& t);

// Majorant Sigma_max(E_i) = max_m lamph[m][i] * maxdens[m] on the table grid (1/cm).
// Takes the role of iniwck (initialize.cu:773-829, 919-966); see DESIGN.md for why it is built on the
// table grid and stored as Sigma (not lambda).
std::vector<float> build_majorant(const Tables& t, const std::vector<float>& maxdens);

std::string join_path(const std::string& base, const std::string& rel);

}  // namespace gpet
