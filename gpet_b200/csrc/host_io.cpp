// Host loaders for gPET's input contract. See host_io.hpp for the reference line citations.
#include "host_io.hpp"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

namespace gpet {

// ------------------------------------------------------------------------------------------------ Scanner
bool Scanner::load(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) return false;
    f.seekg(0, std::ios::end);
    std::streamoff n = f.tellg();
    f.seekg(0, std::ios::beg);
    buf.resize((size_t)n);
    if (n > 0) f.read(&buf[0], n);
    pos = 0;
    ok = true;
    return true;
}

void Scanner::ws() {
    while (pos < buf.size() && isspace((unsigned char)buf[pos])) pos++;
}

std::string Scanner::line(size_t maxlen) {
    size_t start = pos;
    while (pos < buf.size()) {
        if (maxlen && pos - start >= maxlen - 1) break;
        char c = buf[pos++];
        if (c == '\n') break;
    }
    return buf.substr(start, pos - start);
}

bool Scanner::i32(int32_t& v) {
    ws();
    if (eof()) return ok = false;
    char* end = nullptr;
    const char* s = buf.c_str() + pos;
    long r = strtol(s, &end, 10);
    if (end == s) return ok = false;
    v = (int32_t)r;
    pos += (size_t)(end - s);
    ws();
    return true;
}

bool Scanner::u64(uint64_t& v) {
    ws();
    if (eof()) return ok = false;
    char* end = nullptr;
    const char* s = buf.c_str() + pos;
    unsigned long long r = strtoull(s, &end, 10);
    if (end == s) return ok = false;
    v = (uint64_t)r;
    pos += (size_t)(end - s);
    ws();
    return true;
}

bool Scanner::f32(float& v) {
    ws();
    if (eof()) return ok = false;
    char* end = nullptr;
    const char* s = buf.c_str() + pos;
    float r = strtof(s, &end);
    if (end == s) return ok = false;
    v = r;
    pos += (size_t)(end - s);
    ws();
    return true;
}

bool Scanner::f64(double& v) {
    ws();
    if (eof()) return ok = false;
    char* end = nullptr;
    const char* s = buf.c_str() + pos;
    double r = strtod(s, &end);
    if (end == s) return ok = false;
    v = r;
    pos += (size_t)(end - s);
    ws();
    return true;
}

bool Scanner::word(std::string& v) {
    ws();
    if (eof()) return ok = false;
    size_t start = pos;
    while (pos < buf.size() && !isspace((unsigned char)buf[pos])) pos++;
    v = buf.substr(start, pos - start);
    ws();
    return true;
}

void Scanner::ignore_through(size_t n, char delim) {
    size_t cnt = 0;
    while (pos < buf.size() && cnt < n) {
        char c = buf[pos++];
        cnt++;
        if (c == delim) break;
    }
}

bool Scanner::next_is_number() {
    ws();
    if (eof()) return false;
    char c = buf[pos];
    if (isdigit((unsigned char)c)) return true;
    if ((c == '-' || c == '+' || c == '.') && pos + 1 < buf.size()) {
        char d = buf[pos + 1];
        return isdigit((unsigned char)d) || d == '.';
    }
    return false;
}

void Scanner::skip_labels() {
    while (!eof() && !next_is_number()) {
        if (eof()) break;
        line();
    }
}

std::string join_path(const std::string& base, const std::string& rel) {
    if (rel.empty() || rel[0] == '/' || base.empty()) return rel;
    if (base.back() == '/') return base + rel;
    return base + "/" + rel;
}

// ------------------------------------------------------------------------------------------------ input_PET.in
std::string parse_config(const std::string& path, Config& c) {
    Scanner s;
    if (!s.load(path)) return "cannot open config file " + path;
    const size_t L = 200;  // fgets(buffer, 200, ...) in main.cu
    auto ints = [&](int32_t* v, int n) { for (int i = 0; i < n; i++) s.i32(v[i]); };
    auto flts = [&](float* v, int n) { for (int i = 0; i < n; i++) s.f32(v[i]); };
    s.line(L); s.i32(c.device);
    s.line(L); s.f32(c.nonangle);
    s.line(L); ints(c.pdim, 3);
    s.line(L); flts(c.poffset, 3);
    s.line(L); flts(c.psize, 3);
    s.line(L); s.word(c.matfile);
    s.line(L); s.word(c.denfile);
    s.line(L); s.i32(c.nhist);
    s.line(L); s.i32(c.usepsf);
    s.line(L); s.word(c.sourcefile);
    s.line(L); s.i32(c.ptype);
    s.line(L); s.i32(c.useprange);
    s.line(L); s.f32(c.tstart); s.f32(c.tend);
    s.line(L); flts(c.recordsphere, 4);
    s.line(L); s.f32(c.eabsph);
    s.line(L); s.word(c.geofile);
    s.line(L); s.i32(c.nsurface);
    if (!s.ok) return "malformed config file " + path + " (before the surface list)";
    if (c.nsurface < 0 || c.nsurface > GPET_MAX_SURFACES)
        return "config: number of quadric surfaces exceeds MAXSURFACE";  // main.cu:186-190
    c.surface.assign((size_t)10 * c.nsurface, 0.f);
    for (int i = 0; i < 10 * c.nsurface; i++) s.f32(c.surface[i]);
    s.line(L); s.i32(c.rdepth); s.i32(c.rpolicy);
    s.line(L); s.f32(c.Eth);
    s.line(L); s.i32(c.blurpolicy); s.f32(c.Eref); s.f32(c.Rref); s.f32(c.Eslope); s.f32(c.Sblur);
    s.line(L); s.i32(c.dlevel); s.i32(c.dtype); s.f32(c.dtime);
    s.line(L); s.f32(c.Ewinmin); s.f32(c.Ewinmax);
    if (!s.ok) return "malformed config file " + path;
    return "";
}

// ------------------------------------------------------------------------------------------------ .geo
static void rotate_about(const float rot[3], float ang, const float v[3], float out[3]) {
    // component-wise form used by the reference (detector.cu:251-253); exact for axis-aligned axes.
    float ca = cosf(ang), sa = sinf(ang);
    out[0] = (1 - ca) * (v[0] * rot[0]) * rot[0] + ca * v[0] + sa * (rot[1] * v[2] - rot[2] * v[1]);
    out[1] = (1 - ca) * (v[1] * rot[1]) * rot[1] + ca * v[1] + sa * (rot[2] * v[0] - rot[0] * v[2]);
    out[2] = (1 - ca) * (v[2] * rot[2]) * rot[2] + ca * v[2] + sa * (rot[0] * v[1] - rot[1] * v[0]);
}

std::string parse_geometry(const std::string& path, Geometry& g) {
    Scanner s;
    if (!s.load(path)) return "cannot open geometry file " + path;
    const size_t L = 256;
    int32_t count = 0;
    s.line(L); s.i32(count);
    s.line(L); s.f32(g.rot_axis[0]); s.f32(g.rot_axis[1]); s.f32(g.rot_axis[2]);
    s.line(L); s.f32(g.rot_angle_deg);
    s.line(L);
    for (int i = 0; i < 2; i++) { s.i32(g.mat[i]); s.f32(g.dens[i]); }
    s.line(L);
    if (!s.ok || count < 1 || count > 4096) return "malformed geometry file " + path;
    gpet_panel p0;
    memset(&p0, 0, sizeof(p0));
    auto v3 = [&](float& a, float& b, float& c) { s.f32(a); s.f32(b); s.f32(c); };
    s.line(L); s.i32(p0.panel);
    s.line(L); v3(p0.lengthx, p0.lengthy, p0.lengthz);
    s.line(L); v3(p0.MODx, p0.MODy, p0.MODz);
    s.line(L); v3(p0.Mspacex, p0.Mspacey, p0.Mspacez);
    s.line(L); v3(p0.LSOx, p0.LSOy, p0.LSOz);
    s.line(L); v3(p0.spacex, p0.spacey, p0.spacez);
    s.line(L); v3(p0.directionx, p0.directiony, p0.directionz);
    s.line(L); v3(p0.offsetx, p0.offsety, p0.offsetz);
    s.line(L); v3(p0.UniXx, p0.UniXy, p0.UniXz);
    s.line(L); v3(p0.UniYx, p0.UniYy, p0.UniYz);
    s.line(L); v3(p0.UniZx, p0.UniZy, p0.UniZz);
    if (!s.ok) return "malformed geometry file " + path + " (panel block)";
    g.panels.assign((size_t)count, p0);
    const float PI_F = 3.1415926535897932384626433f;
    for (int i = 1; i < count; i++) {
        gpet_panel& p = g.panels[i];
        p.panel = i;
        float ang = g.rot_angle_deg * PI_F / 180.0f * i;
        float v[3], o[3];
        v[0] = p0.offsetx; v[1] = p0.offsety; v[2] = p0.offsetz;
        rotate_about(g.rot_axis, ang, v, o);
        p.offsetx = o[0]; p.offsety = o[1]; p.offsetz = o[2];
        v[0] = p0.UniXx; v[1] = p0.UniXy; v[2] = p0.UniXz;
        rotate_about(g.rot_axis, ang, v, o);
        p.UniXx = o[0]; p.UniXy = o[1]; p.UniXz = o[2];
        v[0] = p0.UniYx; v[1] = p0.UniYy; v[2] = p0.UniYz;
        rotate_about(g.rot_axis, ang, v, o);
        p.UniYx = o[0]; p.UniYy = o[1]; p.UniYz = o[2];
        v[0] = p0.UniZx; v[1] = p0.UniZy; v[2] = p0.UniZz;
        rotate_about(g.rot_axis, ang, v, o);
        p.UniZx = o[0]; p.UniZy = o[1]; p.UniZz = o[2];
    }
    // module / crystal counts from panel 0 only (initialize.cu:1074-1086); int *= float as in the reference
    const gpet_panel& q = g.panels[0];
    int Mn = (int)(floorf(q.lengthy / (q.MODy + q.Mspacey)) + 1);
    int Ln = (int)(floorf(q.MODy / (q.LSOy + q.spacey)) + 1);
    g.moduleNy = Mn;
    g.crystalNy = Ln;
    Mn = (int)(Mn * (floorf(q.lengthz / (q.MODz + q.Mspacez)) + 1));
    Ln = (int)(Ln * (floorf(q.MODz / (q.LSOz + q.spacez)) + 1));
    g.moduleN = Mn;
    g.crystalN = Ln;
    return "";
}

// ------------------------------------------------------------------------------------------------ isotopes / sources
std::string parse_isotopes(const std::string& path, Isotopes& iso) {
    Scanner s;
    if (!s.load(path)) return "cannot open isotope file " + path;
    int32_t n = 0;
    s.i32(n);
    s.ignore_through(512, '#');
    if (!s.ok || n < 1 || n > GPET_MAX_ISOTOPES) return "malformed isotope file " + path;
    iso.halftime.assign(n, 0.f);
    iso.ratio.assign(n, 0.f);
    iso.coef.assign((size_t)8 * n, 0.f);
    for (int i = 0; i < n; i++) {
        s.f32(iso.halftime[i]);
        s.f32(iso.ratio[i]);
        for (int j = 0; j < 8; j++) s.f32(iso.coef[8 * i + j]);
    }
    if (!s.ok) return "malformed isotope file " + path;
    return "";
}

std::string parse_sources(const std::string& path, Sources& src) {
    Scanner s;
    if (!s.load(path)) return "cannot open source file " + path;
    int32_t n = 0;
    s.i32(n);
    s.ignore_through(512, '#');
    if (!s.ok || n < 1 || n > GPET_MAX_SOURCES) return "malformed source file " + path;
    src.natom.assign(n, 0);
    src.type.assign(n, 0);
    src.shape.assign(n, 0);
    src.coeff.assign((size_t)6 * n, 0.f);
    for (int i = 0; i < n; i++) {
        s.u64(src.natom[i]);
        s.i32(src.type[i]);
        s.i32(src.shape[i]);
        for (int j = 0; j < 6; j++) s.f32(src.coeff[6 * i + j]);
    }
    if (!s.ok) return "malformed source file " + path;
    return "";
}

// ------------------------------------------------------------------------------------------------ phantom / psf
static std::string read_binary(const std::string& path, void* dst, size_t bytes) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return "cannot open " + path;
    size_t got = fread(dst, 1, bytes, f);
    fclose(f);
    if (got != bytes) return "short read on " + path;
    return "";
}

std::string load_phantom(const std::string& matfile, const std::string& denfile, const int32_t dim[3],
                         const float offset[3], const float size[3], Phantom& ph) {
    for (int i = 0; i < 3; i++) {
        if (dim[i] < 1) return "phantom dimension must be positive";
        ph.dim[i] = dim[i];
        ph.offset[i] = offset[i];
        ph.size[i] = size[i];
        ph.d[i] = size[i] / dim[i];  // initialize.cu:68-70
    }
    size_t n = ph.nvox();
    ph.mat.resize(n);
    ph.dens.resize(n);
    std::string e = read_binary(matfile, ph.mat.data(), n * sizeof(int32_t));
    if (!e.empty()) return e;
    return read_binary(denfile, ph.dens.data(), n * sizeof(float));
}

std::string load_psf(const std::string& path, int64_t max_particles, int ptype, Psf& psf) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return "cannot open psf file " + path;
    fseek(f, 0, SEEK_END);
    int64_t bytes = ftell(f);
    fseek(f, 0, SEEK_SET);
    int64_t n = bytes / 64;
    if (max_particles > 0 && max_particles < n) n = max_particles;  // initialize.cu:88-92
    psf.p.resize((size_t)n);
    psf.ptype = ptype;
    std::vector<double> rec((size_t)8 * 4096);
    int64_t done = 0;
    while (done < n) {
        int64_t chunk = std::min<int64_t>(4096, n - done);
        if (fread(rec.data(), 64, (size_t)chunk, f) != (size_t)chunk) { fclose(f); return "short read on " + path; }
        for (int64_t i = 0; i < chunk; i++) {
            const double* d = &rec[(size_t)8 * i];
            gpet_photon& p = psf.p[(size_t)(done + i)];
            p.x = (float)d[0]; p.y = (float)d[1]; p.z = (float)d[2];
            p.t = d[3];
            p.vx = (float)d[4]; p.vy = (float)d[5]; p.vz = (float)d[6];
            p.E = (float)d[7];
            p.nscat = 0;
            p.eventid = (int32_t)(done + i);  // initialize.cu:103
            p.parn = (int32_t)(done + i);
        }
        done += chunk;
    }
    fclose(f);
    return "";
}

// ------------------------------------------------------------------------------------------------ tables
static std::string read_matter(const std::string& path, Tables& t) {
    Scanner s;
    if (!s.load(path)) return "cannot open " + path;
    float tmp;
    s.skip_labels(); s.f32(t.eminph); s.f32(tmp); s.f32(t.emax);
    s.skip_labels(); s.f32(tmp); s.f32(tmp);
    s.skip_labels(); s.f32(tmp); s.f32(tmp); s.f32(tmp);
    s.skip_labels(); s.i32(t.nmat);
    if (!s.ok || t.nmat < 1 || t.nmat > GPET_MAX_MATERIALS) return "malformed material file " + path;
    t.names.clear();
    t.refdens.assign(t.nmat, 0.f);
    for (int m = 0; m < t.nmat; m++) {
        std::string name;
        while (!s.eof()) {
            std::string ln = s.line();
            size_t k = ln.find("MATERIAL:");
            if (k != std::string::npos) {
                name = ln.substr(k + 9);
                while (!name.empty() && isspace((unsigned char)name.back())) name.pop_back();
                while (!name.empty() && isspace((unsigned char)name.front())) name.erase(name.begin());
                break;
            }
        }
        t.names.push_back(name);
        int32_t nelem = 0;
        s.skip_labels(); s.f32(t.refdens[m]);
        s.skip_labels(); s.i32(nelem);
        for (int j = 0; j < nelem; j++) { int32_t z; s.i32(z); s.f32(tmp); }
        s.skip_labels(); s.f32(tmp); s.f32(tmp); s.f32(tmp);
        s.skip_labels(); s.f32(tmp);
        s.skip_labels(); s.f32(tmp); s.f32(tmp);
        if (!s.ok) return "malformed material block in " + path;
    }
    return "";
}

static std::string read_1d(const std::string& path, int nmat, std::vector<float>& energy, std::vector<float>& val,
                           int& nen) {
    Scanner s;
    if (!s.load(path)) return "cannot open " + path;
    for (int m = 0; m < nmat; m++) {
        int32_t nd = 0;
        float tmp;
        s.skip_labels();
        s.i32(nd); s.f32(tmp); s.f32(tmp); s.f32(tmp); s.f32(tmp);
        if (!s.ok || nd < 2) return "malformed table header in " + path;
        if (m == 0) {
            if (nen == 0) nen = nd;
            val.assign((size_t)nmat * nd, 0.f);
        }
        if (nd != nen) return "table dimensions differ between materials/files in " + path;
        std::vector<float> e((size_t)nd);
        s.skip_labels();
        for (int i = 0; i < nd; i++) { s.f32(e[i]); s.f32(val[(size_t)m * nd + i]); }
        if (!s.ok) return "short table in " + path;
        if (energy.empty()) energy = e;
        else if (fabsf(energy[0] - e[0]) > 1e-3f * fabsf(e[0]) || fabsf(energy[nd - 1] - e[nd - 1]) > 1e-6f * e[nd - 1])
            return "energy grids differ between tables (" + path + ")";
    }
    return "";
}

static std::string read_surface(const std::string& path, int nmat, int& ncp, int& ne, float& dcp, float& de,
                                std::vector<float>& surf) {
    Scanner s;
    if (!s.load(path)) return "cannot open " + path;
    for (int m = 0; m < nmat; m++) {
        int32_t nd = 0;
        float tmp;
        s.skip_labels();
        s.i32(nd); s.f32(tmp); s.f32(tmp); s.f32(tmp);
        s.skip_labels();
        for (int i = 0; i < 3 * nd; i++) s.f32(tmp);  // q, ln q, S(q)|F(q): not used by the transport
        s.skip_labels();
        int32_t ncp_m = 0, ne_m = 0;
        float dcp_m = 0.f, de_m = 0.f;
        s.i32(ncp_m); s.f32(tmp); s.f32(tmp); s.f32(dcp_m); s.i32(ne_m); s.f32(tmp); s.f32(tmp); s.f32(de_m);
        if (!s.ok || ncp_m < 2 || ne_m < 2) return "malformed surface header in " + path;
        if (m == 0) {
            ncp = ncp_m; ne = ne_m; dcp = dcp_m; de = de_m;
            surf.assign((size_t)nmat * ncp * ne, 0.f);
        } else if (ncp_m != ncp || ne_m != ne) {
            return "surface dimensions differ between materials in " + path;
        }
        for (int i = 0; i < ncp; i++) s.f32(tmp);
        for (int i = 0; i < ne; i++) s.f32(tmp);
        float* dst = &surf[(size_t)m * ncp * ne];
        for (int i = 0; i < ncp * ne; i++) s.f32(dst[i]);  // [icp][ie], initialize.cu:534-543
        if (!s.ok) return "short surface in " + path;
    }
    return "";
}

std::string load_tables_ascii(const std::string& prefix, Tables& t) {
    t = Tables();
    std::string e = read_matter(prefix + ".matter", t);
    if (!e.empty()) return e;
    int nen = 0;
    if (!(e = read_1d(prefix + ".lamph", t.nmat, t.energy, t.lamph, nen)).empty()) return e;
    if (!(e = read_1d(prefix + ".compt", t.nmat, t.energy, t.compt, nen)).empty()) return e;
    if (!(e = read_1d(prefix + ".phote", t.nmat, t.energy, t.phote, nen)).empty()) return e;
    if (!(e = read_1d(prefix + ".rayle", t.nmat, t.energy, t.rayle, nen)).empty()) return e;
    t.nen = nen;
    if (!(e = read_surface(prefix + ".cmpsf", t.nmat, t.cm_ncp, t.cm_ne, t.cm_dcp, t.cm_de, t.cmpsf)).empty()) return e;
    if (!(e = read_surface(prefix + ".rayff", t.nmat, t.rl_ncp, t.rl_ne, t.rl_dcp, t.rl_de, t.rayff)).empty()) return e;
    return "";
}

static const char kMagic[8] = {'G', 'P', 'E', 'T', 'T', 'A', 'B', '1'};

std::string save_tables_packed(const std::string& path, const Tables& t) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return "cannot write " + path;
    auto wi = [&](int32_t v) { fwrite(&v, 4, 1, f); };
    auto wf = [&](float v) { fwrite(&v, 4, 1, f); };
    auto wv = [&](const std::vector<float>& v) { if (!v.empty()) fwrite(v.data(), 4, v.size(), f); };
    fwrite(kMagic, 1, 8, f);
    wi(t.nmat); wi(t.nen); wf(t.eminph); wf(t.emax);
    wi(t.cm_ncp); wi(t.cm_ne); wf(t.cm_dcp); wf(t.cm_de);
    wi(t.rl_ncp); wi(t.rl_ne); wf(t.rl_dcp); wf(t.rl_de);
    for (int m = 0; m < t.nmat; m++) {
        char name[32];
        memset(name, 0, sizeof(name));
        strncpy(name, t.names[m].c_str(), 31);
        fwrite(name, 1, 32, f);
        wf(t.refdens[m]);
    }
    wv(t.energy); wv(t.lamph); wv(t.compt); wv(t.phote); wv(t.rayle); wv(t.cmpsf); wv(t.rayff);
    fclose(f);
    return "";
}

std::string load_tables_packed(const std::string& path, Tables& t) {
    t = Tables();
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return "cannot open " + path;
    bool good = true;
    auto ri = [&](int32_t& v) { good &= fread(&v, 4, 1, f) == 1; };
    auto rf = [&](float& v) { good &= fread(&v, 4, 1, f) == 1; };
    auto rv = [&](std::vector<float>& v, size_t n) { v.resize(n); if (n) good &= fread(v.data(), 4, n, f) == n; };
    char magic[8];
    good &= fread(magic, 1, 8, f) == 8;
    if (!good || memcmp(magic, kMagic, 8) != 0) { fclose(f); return "not a packed table file: " + path; }
    ri(t.nmat); ri(t.nen); rf(t.eminph); rf(t.emax);
    ri(t.cm_ncp); ri(t.cm_ne); rf(t.cm_dcp); rf(t.cm_de);
    ri(t.rl_ncp); ri(t.rl_ne); rf(t.rl_dcp); rf(t.rl_de);
    if (!good || t.nmat < 1 || t.nmat > GPET_MAX_MATERIALS || t.nen < 2 || t.nen > (1 << 20)) {
        fclose(f);
        return "corrupt packed table header: " + path;
    }
    t.refdens.resize(t.nmat);
    for (int m = 0; m < t.nmat; m++) {
        char name[33];
        memset(name, 0, sizeof(name));
        good &= fread(name, 1, 32, f) == 32;
        t.names.push_back(name);
        rf(t.refdens[m]);
    }
    size_t n1 = (size_t)t.nmat * t.nen;
    rv(t.energy, t.nen); rv(t.lamph, n1); rv(t.compt, n1); rv(t.phote, n1); rv(t.rayle, n1);
    rv(t.cmpsf, (size_t)t.nmat * t.cm_ncp * t.cm_ne);
    rv(t.rayff, (size_t)t.nmat * t.rl_ncp * t.rl_ne);
    fclose(f);
    if (!good) return "short packed table file: " + path;
    return "";
}

std::vector<float> build_majorant(const Tables& t, const std::vector<float>& maxdens) {
    std::vector<float> maj((size_t)t.nen, 0.f);
    for (int i = 0; i < t.nen; i++) {
        float ymax = 0.f;
        for (int m = 0; m < t.nmat && m < (int)maxdens.size(); m++) {
            float y = t.lamph[(size_t)m * t.nen + i] * maxdens[m];
            if (y > ymax) ymax = y;
        }
        maj[i] = ymax;
    }
    return maj;
}

}  // namespace gpet
