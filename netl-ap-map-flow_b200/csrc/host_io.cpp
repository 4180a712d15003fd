// host_io.cpp -- the file-format side of the drop-in boundary: apm_lcl_run_file().
//
// Host C++ equivalents of the reference's I/O layer, written against its behaviour, not its code:
//   LNPROC tokenizer            apmapflow/src/STRING_MODULE.F:114-239
//   OPEN_IO / OPEN_FILE         apmapflow/src/IO_MODULE.F:51-344       (file keywords, overwrite rules)
//   READ_MASTER                 apmapflow/src/APERTURE_MAP_FLOW.F:191-405 (keyword chain, defaults :88-126)
//   unit tables                 apmapflow/src/UNIT_CONVERSION_MODULE.F:82-360
//   READ_DATA_MAP               apmapflow/src/MAP_MODULE.F:18-85
//   MAP_STATS                   apmapflow/src/MAP_MODULE.F:182-303
//   CALC_PERCENTILE             apmapflow/src/OUTPUT_MODULE.F:53-124
//   OUTPUT_DATA / INIT_OUTPUT / WRITE_STATS_FILE / WRITE_VTK_FILE   APERTURE_MAP_FLOW.F:773-1291
// The numeric hot path is not here: it is one call to apm_lcl_solve() (api.cu).
#include <algorithm>
#include <cctype>
#include <cerrno>
#include <charconv>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <mutex>
#include <sstream>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <sys/stat.h>
#include <unistd.h>

#include "../../include/apm_lcl.h"
#include "fortran_format.hpp"

namespace {

using ffmt::A;
using ffmt::E;
using ffmt::F;
using ffmt::G;
using ffmt::I;
using ffmt::append_E;
using ffmt::append_G;

constexpr int kMaxFld = 100, kMaxLen = 256;   // STRING_MODULE.F:21-22

std::string upper(std::string s)
{
    for (auto &c : s)
        if (c >= 'a' && c <= 'z') c = (char)(c - 32);
    return s;
}
std::string lower(std::string s)
{
    for (auto &c : s)
        if (c >= 'A' && c <= 'Z') c = (char)(c + 32);
    return s;
}
bool has(const std::string &s, const char *sub) { return s.find(sub) != std::string::npos; }
std::string rtrim(std::string s)
{
    while (!s.empty() && (s.back() == ' ' || s.back() == '\r')) s.pop_back();
    return s;
}
// CHARACTER(n) variable written with An / A: left justified, blank padded or cut
std::string fixed(const std::string &s, size_t n)
{
    std::string r = s.substr(0, std::min(n, s.size()));
    r.resize(n, ' ');
    return r;
}

// ------------------------------------------------------------------------------------------
// Host-side parallelism for the ASCII pipeline (once the solve takes milliseconds,
// text I/O is the run time).  Rows are formatted by a pool of threads in waves and written in order, so
// the files are byte-identical to the sequential writer.
unsigned io_threads()
{
    if (const char *e = getenv("APM_LCL_IO_THREADS")) {   // 1 = the sequential writers and reader
        const int k = atoi(e);
        if (k >= 1) return k > 256 ? 256u : (unsigned)k;
    }
    unsigned n = std::thread::hardware_concurrency();
    if (n == 0) n = 4;
    return n > 32 ? 32 : n;
}

// APM_LCL_IO_TIMING=1: wall time of every stage of the file pipeline on stderr
struct StageTimer {
    bool on = getenv("APM_LCL_IO_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char *what)
    {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[apm_lcl io] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - t).count());
        t = now;
    }
};

template <class Fn>   // fn(row, std::string &out) appends the text of one row
void write_rows_parallel(FILE *f, long long nrows, Fn fn)
{
    const unsigned T = io_threads();
    const long long per_task = 32;
    if (nrows <= per_task * 2 || T == 1) {
        std::string buf;
        for (long long r = 0; r < nrows; r++) {
            fn(r, buf);
            if (buf.size() > (1u << 20)) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); }
        }
        fwrite(buf.data(), 1, buf.size(), f);
        return;
    }
    // T formatting threads take chunks of per_task rows from a counter and fill a ring of buffers; this
    // thread writes the chunks out in order while later ones are still being formatted.
    const long long nchunks = (nrows + per_task - 1) / per_task;
    const long long window = 2LL * T;
    std::vector<std::string> bufs((size_t)window);
    std::vector<long long> done((size_t)window, -1);   // chunk whose text the slot holds
    std::mutex m;
    std::condition_variable cv;
    long long next = 0, written = 0;
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < T; t++)
        pool.emplace_back([&]() {
            // A thread formats into a string of its own and swaps it into the ring when the chunk is complete.  Formatting
            // straight into the ring's strings cost 2.7x: the string headers of neighbouring slots share cache lines, and
            // every append updates a size field (measured, 4096^2 CSV map on 8 cores: 0.40 s against 0.15 s).
            std::string buf;
            for (;;) {
                long long ch;
                {
                    std::unique_lock<std::mutex> lock(m);
                    ch = next++;
                    if (ch >= nchunks) return;
                }
                buf.clear();
                const long long b = ch * per_task, e = std::min(nrows, b + per_task);
                for (long long r = b; r < e; r++) fn(r, buf);
                {
                    std::unique_lock<std::mutex> lock(m);
                    cv.wait(lock, [&]() { return ch < written + window; });   // its slot has been written out
                    bufs[(size_t)(ch % window)].swap(buf);
                    done[(size_t)(ch % window)] = ch;
                }
                cv.notify_all();
            }
        });
    for (long long ch = 0; ch < nchunks; ch++) {
        {
            std::unique_lock<std::mutex> lock(m);
            cv.wait(lock, [&]() { return done[(size_t)(ch % window)] == ch; });
        }
        const std::string &buf = bufs[(size_t)(ch % window)];
        fwrite(buf.data(), 1, buf.size(), f);
        {
            std::lock_guard<std::mutex> lock(m);
            written = ch + 1;
        }
        cv.notify_all();
    }
    for (auto &th : pool) th.join();
}

template <class Fn>   // fn(begin, end, slot) over [0, n) split into `slots` contiguous parts
void parallel_ranges(long long n, unsigned slots, Fn fn)
{
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < slots; t++) {
        const long long b = n * t / slots, e = n * (t + 1) / slots;
        pool.emplace_back([&, b, e, t]() { fn(b, e, t); });
    }
    for (auto &th : pool) th.join();
}

// ------------------------------------------------------------------------------------------
// LNPROC: splits a record into fields.  ';' starts a comment, blanks / ',' / TAB separate fields,
// '=' separates without creating empty fields, double quotes protect everything in between; the
// QUOTED flag is SAVEd in the reference, so an unterminated quote leaks into the next record.
struct LineReader {
    std::istream &in;
    bool quoted = false;
    explicit LineReader(std::istream &s) : in(s) {}
    // returns number of fields (0 = end of file); skips records without fields
    int next(std::vector<std::string> &fld, std::string &record, int maxfld, const char *kase = "UPPER")
    {
        std::string line;
        while (std::getline(in, line)) {
            line = rtrim(line);
            if (has(upper(kase), "UP")) line = upper(line);
            else if (has(upper(kase), "LO")) line = lower(line);
            record = line;
            fld.clear();
            int leng = 0;   // >0 inside a field, 0 after a ',' / TAB, -1 after a blank or '='
            bool stop = false;
            const size_t n = line.size();
            for (size_t i = 0; i <= n && !stop; i++) {
                const char c = i < n ? line[i] : ' ';
                auto put = [&](char ch) {
                    if (leng <= 0) {
                        if ((int)fld.size() == maxfld) { stop = true; return; }
                        fld.emplace_back(1, ch);
                        leng = 1;
                    } else if (leng < kMaxLen) {
                        fld.back().push_back(ch);
                        leng++;
                    }
                };
                if (quoted) {
                    if (c == '"') quoted = false;
                    else put(c);
                } else if (c == ';') {
                    break;
                } else if (c == '"') {
                    quoted = true;
                } else if (c == ' ' || c == '=') {
                    if (leng > 0) leng = -1;
                } else if (c == ',' || c == '\t') {
                    if (leng > 0) leng = 0;
                    else if (leng == 0) {
                        if ((int)fld.size() == maxfld) { stop = true; break; }
                        fld.emplace_back(" ");
                    } else leng = 0;
                } else {
                    put(c);
                }
            }
            if (stop) return (int)fld.size();   // MAXFLD reached: the reference returns what it has
            if (!fld.empty()) return (int)fld.size();
        }
        fld.clear();
        return 0;
    }
};

// ------------------------------------------------------------------------------------------
// unit tables (substring matched, in the reference's order)
double conv_dist(const std::string &u, double vox)
{
    if (has(u, "MICR")) return 1.0e-6;
    if (has(u, "MM")) return 1.0e-3;
    if (has(u, "CM")) return 0.01;
    if (has(u, "IN")) return 0.0254;
    if (has(u, "FT") || has(u, "FEET")) return 0.3048;
    if (has(u, "ML")) return 0.01;
    if (has(u, "VOX")) return vox;
    if (has(u, "M")) return 1.0;
    return -1.0;
}
double conv_mass(const std::string &u)
{
    if (has(u, "SLUG")) return 14.59390;
    if (has(u, "LB")) return 0.45359237;
    if (has(u, "KG")) return 1.0;
    if (has(u, "G")) return 1.0e-3;
    return -1.0;
}
double conv_time(const std::string &u)
{
    if (has(u, "DAY")) return 86400.0;
    if (has(u, "HOUR") || has(u, "HR")) return 3600.0;
    if (has(u, "MIN")) return 60.0;
    if (has(u, "SEC")) return 1.0;
    return -1.0;
}
double conv_pres(const std::string &u)
{
    if (has(u, "PSI")) return 6894.757293178;
    if (has(u, "BAR")) return 1.0e5;
    if (has(u, "ATM")) return 101325.0;
    if (has(u, "KPA")) return 1.0e3;
    if (has(u, "PA")) return 1.0;
    return -1.0;
}
double conv_visc(const std::string &u)
{
    if (has(u, "CP") || has(u, "CENT")) return 1.0e-3;
    if (has(u, "PA") || has(u, "SEC")) return 1.0;
    return -1.0;
}

struct Units {
    double vox = 1.0;
    std::function<void(const std::string &)> say;
    static const char *si_name(const std::string &t)
    {
        if (has(t, "DIST")) return "M";
        if (has(t, "MASS")) return "KG";
        if (has(t, "TIME")) return "SEC";
        if (has(t, "DENS")) return "KG/M^3";
        if (has(t, "FLOW") || has(t, "RATE")) return "M^3/SEC";
        if (has(t, "PRES")) return "PA";
        if (has(t, "VISC")) return "PA*SEC";
        return nullptr;
    }
    double to_si(const std::string &type, const std::string &unit_in) const
    {
        const std::string u = upper(unit_in);
        if (has(type, "DIST")) return conv_dist(u, vox);
        if (has(type, "MASS")) return conv_mass(u);
        if (has(type, "TIME")) return conv_time(u);
        if (has(type, "DENS") || has(type, "FLOW") || has(type, "RATE")) {
            const size_t k = u.find('/');
            const std::string a = k == std::string::npos ? std::string() : u.substr(0, k);
            const std::string b = k == std::string::npos ? u : u.substr(k + 1);
            if (has(type, "DENS")) {
                const double m = conv_mass(a), d = conv_dist(b, vox);
                if (m < 0) say(" NO DENSITY, MASS CONVERSION FOUND FOR UNIT: " + a);
                if (d < 0) say(" NO DENSITY, DISTANCE CONVERSION FOUND FOR UNIT: " + b);
                return m / (d * d * d);
            }
            const double d = conv_dist(a, vox), t = conv_time(b);
            if (d < 0) say(" NO RATE, DISTANCE CONVERSION FOUND FOR UNIT: " + a);
            if (t < 0) say(" NO RATE, TIME CONVERSION FOUND FOR UNIT: " + b);
            return d * d * d / t;
        }
        if (has(type, "PRES")) return conv_pres(u);
        if (has(type, "VISC")) return conv_visc(u);
        return -1.0;
    }
    // CALC_CONV_FACT: unit_out == "SI" is replaced by the SI unit name
    bool factor(const std::string &type, const std::string &unit_in, std::string &unit_out, double &f) const
    {
        const char *si = si_name(type);
        if (!si) {
            say("  NO CONVERTER FOUND FOR UNIT TYPE: " + type);
            return false;
        }
        const double a = to_si(type, unit_in);
        if (rtrim(unit_out) == "SI") {
            if (a < 0) { say(" NO " + type + " CONVERSION FOUND FOR " + upper(unit_in) + " TO " + unit_out); return false; }
            f = a;
            unit_out = si;
            return true;
        }
        const double b = to_si(type, unit_out);
        if (a < 0 || b < 0) { say(" NO " + type + " CONVERSION FOUND FOR " + upper(unit_in) + " TO " + unit_out); return false; }
        f = a / b;
        return true;
    }
};

// Fortran "(F100.0)" / list-directed read of one field
bool read_real(const std::string &s, double &v)
{
    std::string t = rtrim(s);
    size_t b = 0;
    while (b < t.size() && t[b] == ' ') b++;
    t = t.substr(b);
    if (t.empty()) { v = 0.0; return true; }   // blank field reads as zero under F editing
    for (auto &c : t)
        if (c == 'd' || c == 'D') c = 'E';
    errno = 0;
    char *end = nullptr;
    v = strtod(t.c_str(), &end);
    if (end == t.c_str() || *end != '\0') return false;
    for (char c : t)
        if (std::isalpha((unsigned char)c) && c != 'E' && c != 'e') return false;   // no "inf"/"nan"/hex
    return true;
}

// ------------------------------------------------------------------------------------------
struct Run {
    // messages go to the caller's buffer (stdout of the reference) and, once open, to the log file
    std::string out;
    FILE *flog = nullptr;
    bool open_io_failed = false;   // OPEN_IO gave up: every file it had created is removed again (IO_MODULE.F:219-226)
    void to_stdout(const std::string &s) { out += s; out += '\n'; }
    void message(const std::string &s)
    {
        const std::string t = rtrim(s);
        to_stdout(t);
        if (flog) fprintf(flog, "%s\n", t.c_str());
    }
    void blank() { message(" "); }

    // file names
    std::string inp, apm_file;
    std::string sum_file = "./fracture-log.txt", stat_file = "./fracture-stat.csv", map_file = "./fracture-aperture.csv",
                flow_file = "./fracture-flow.csv", press_file = "./fracture-press.csv", vtk_file = "./fracture-data.vtk";
    std::string flox, floz, flom, statc, staty;
    bool overwrite = false;
    std::vector<std::pair<std::string, FILE *>> opened;
    FILE *fstatc = nullptr, *fstaty = nullptr, *fmap = nullptr, *fflox = nullptr, *ffloz = nullptr, *fflom = nullptr,
         *fpres = nullptr, *fvtk = nullptr;

    // model state (APM_MODULE.F:74-98)
    apm_lcl_params p{};
    Units units;
    std::string outlet_side = "TOP";
    std::string unit_in[8], unit_si[8], unit_out[8];
    double pres_conv = 1, dim_conv = 1, time_conv = 1, rate_conv = 1, visc_conv = 1, rho_conv = 1;
    bool percentile = false;
    std::vector<double> perc = {1, 5, 10, 25, 50, 75, 90, 95, 99};
    int maxdim = 2000;
    int nx = 0, nz = 0;
    std::vector<double> ap;
    double aper_rms = 0, aper_dev = 0, x_corr = 0, z_corr = 0;

    bool open_out(FILE *&f, const std::string &name)
    {
        struct stat st;
        if (!overwrite && stat(name.c_str(), &st) == 0) {   // STATUS='NEW'
            to_stdout(" ");
            to_stdout("FILE: " + name.substr(0, 120) + " ALREADY EXISTS");
            to_stdout(" SPECIFY \"OVERWRITE\" IN THE INITIALIZATION FILE OR USE A DIFFERENT NAME FOR OUTPUT FILE: " + name);
            return false;
        }
        f = fopen(name.c_str(), "w");
        if (!f) {
            to_stdout(" ");
            to_stdout("IO ERROR: " + std::to_string(errno) + " ATTEMPTING TO OPEN: " + name.substr(0, 120));
            return false;
        }
        opened.emplace_back(name, f);
        return true;
    }
    void close_all(bool remove_files)
    {
        for (auto &o : opened) {
            if (o.second) fclose(o.second);
            if (remove_files) unlink(o.first.c_str());   // IO_MODULE.F:219-226
        }
        opened.clear();
        flog = nullptr;
    }
};

std::string insert_before_ext(const std::string &name, const std::string &tag)
{
    const size_t k = name.rfind('.');
    if (k == std::string::npos) return name + tag;   // (the reference would misbehave; be kind)
    return name.substr(0, k) + tag + name.substr(k);
}

void insufficient(Run &r, const std::string &where, const std::string &record)
{
    r.blank();
    r.message("     **** ERROR WHILE READING " + where + " DATA ****");
    r.message("     **** THE FOLLOWING RECORD CONTAINS INSUFFICIENT DATA ****");
    r.message(record);
}
void field_error(Run &r, const std::string &where, int ifld, const std::string &record)
{
    r.blank();
    r.message("     **** ERROR WHILE READING " + where + " DATA ****");
    r.message("     **** FIELD " + I(4, 1, ifld) + " ON THE FOLLOWING RECORD IS IN ERROR ****");
    r.message(record);
}

// OPEN_IO, IO_MODULE.F:51-234.  Leaves `in` positioned at the first non-file keyword record.
bool open_io(Run &r, std::ifstream &in, std::streampos &resume)
{
    r.to_stdout("");
    r.to_stdout(" USING SUPPLIED INPUT FILE " + r.inp);
    in.open(r.inp);
    if (!in) {
        r.to_stdout("IO ERROR: " + std::to_string(errno) + " ATTEMPTING TO OPEN: " + r.inp);
        return false;
    }
    r.to_stdout("");
    r.to_stdout(" READING INITIALIZATION FILE");
    LineReader lr(in);
    std::vector<std::string> f;
    std::string rec;
    bool apm = false;
    for (;;) {
        const std::streampos before = in.tellg();
        const int n = lr.next(f, rec, kMaxFld, "SAME");
        if (n == 0) { resume = in.tellg(); in.clear(); break; }
        std::string *target = nullptr;
        if (has(f[0], "APER-MAP")) { target = &r.apm_file; apm = true; }
        else if (has(f[0], "SUMMARY-FILE")) target = &r.sum_file;
        else if (has(f[0], "STAT-FILE")) target = &r.stat_file;
        else if (has(f[0], "APER-FILE")) target = &r.map_file;
        else if (has(f[0], "FLOW-FILE")) target = &r.flow_file;
        else if (has(f[0], "PRESS-FILE")) target = &r.press_file;
        else if (has(f[0], "VTK-FILE")) target = &r.vtk_file;
        else if (has(f[0], "OVERWRITE")) { r.overwrite = true; continue; }
        else { resume = before; break; }   // BACKSPACE
        if (n < 2) {
            insufficient(r, "FILE NAME", rec);
            return false;
        }
        *target = f[1];
        if (target == &r.vtk_file) {
            const size_t k = r.vtk_file.rfind('.');
            if (k != std::string::npos) r.vtk_file = r.vtk_file.substr(0, k) + lower(r.vtk_file.substr(k));
        }
    }
    if (!apm) {
        r.message("FATAL ERROR - NO APERTURE MAP INPUT FILE PROVIDED");
        return false;
    }
    r.flox = insert_before_ext(r.flow_file, "-x");
    r.floz = insert_before_ext(r.flow_file, "-z");
    r.flom = insert_before_ext(r.flow_file, "-m");
    {
        const size_t k = r.stat_file.rfind('.');
        const std::string root = k == std::string::npos ? r.stat_file + "." : r.stat_file.substr(0, k + 1);
        r.statc = root + "csv";
        r.staty = root + "yaml";
    }
    bool ok = true;
    FILE *flog = nullptr;
    ok &= r.open_out(flog, r.sum_file);
    r.flog = flog;
    ok &= r.open_out(r.fstatc, r.statc);
    ok &= r.open_out(r.fstaty, r.staty);
    ok &= r.open_out(r.fmap, r.map_file);
    ok &= r.open_out(r.fflox, r.flox);
    ok &= r.open_out(r.ffloz, r.floz);
    ok &= r.open_out(r.fflom, r.flom);
    ok &= r.open_out(r.fpres, r.press_file);
    ok &= r.open_out(r.fvtk, r.vtk_file);
    if (!ok) return false;
    r.message("     IO INITIALIZED");
    return true;
}

// READ_MASTER, APERTURE_MAP_FLOW.F:191-405
bool read_master(Run &r, std::ifstream &in)
{
    LineReader lr(in);
    std::vector<std::string> f;
    std::string rec;
    bool output_unit = false;
    int n;
    auto bad_field = [&](int ifld) { field_error(r, "INITIALIZATION", ifld, rec); return false; };
    auto too_few = [&]() { insufficient(r, "INITIALIZATION", rec); return false; };
    auto value_unit = [&](double &val, int slot, const char *type) -> int {   // fields 2 (value) and 3 (unit)
        if (n < 3) return -1;
        if (!read_real(f[1], val)) return 2;
        r.unit_in[slot] = f[2];
        double fac;
        std::string si = "SI";
        if (r.units.factor(type, r.unit_in[slot], si, fac)) val *= fac;
        else return -2;
        r.unit_si[slot] = si;
        return 0;
    };
    while ((n = lr.next(f, rec, kMaxFld, "SAME")) > 0) {
        const std::string key = upper(f[0]);
        int st = 0;
        if (has(key, "FRAC-PRESS")) {
            r.message("NOTICE - KEYWORD 'FRAC-PRESS' IS DEPRECATED");
            st = value_unit(r.p.inlet_pressure, 1, "PRES");
        } else if (has(key, "OUTFLOW-SIDE")) {
            r.message("NOTICE - KEYWORD 'OUTFLOW-SIDE' IS DEPRECATED");
            if (n < 2) return too_few();
            r.outlet_side = f[1];
        } else if (has(key, "INLET-PRESS")) {
            st = value_unit(r.p.inlet_pressure, 1, "PRES");
        } else if (has(key, "OUTLET-PRES")) {
            st = value_unit(r.p.outlet_pressure, 1, "PRES");
        } else if (has(key, "OUTLET-RATE")) {
            st = value_unit(r.p.outlet_rate, 5, "RATE");
        } else if (has(key, "INLET-RATE")) {
            st = value_unit(r.p.inlet_rate, 5, "RATE");
        } else if (has(key, "FLUID-DENSITY")) {
            if (n < 2) return too_few();
            if (!read_real(f[1], r.p.density)) return bad_field(2);
            r.unit_in[7] = n >= 3 ? f[2] : "";
            double fac;
            std::string si = "SI";
            if (r.units.factor("DENS", r.unit_in[7], si, fac)) r.p.density *= fac;
            else st = -2;
        } else if (has(key, "FLUID-VISCOSITY")) {
            if (n < 2) return too_few();
            if (!read_real(f[1], r.p.viscosity)) return bad_field(2);
            r.unit_in[6] = n >= 3 ? f[2] : "";
            double fac;
            std::string si = "SI";
            if (r.units.factor("VISC", r.unit_in[6], si, fac)) { r.visc_conv = fac; r.p.viscosity *= fac; }
            else st = -2;
        } else if (has(key, "OUTLET-SIDE")) {
            if (n < 2) return too_few();
            r.outlet_side = upper(f[1]);
        } else if (has(key, "VOXEL")) {
            if (n < 4) return too_few();
            if (!read_real(f[2], r.units.vox)) return bad_field(3);
            double fac;
            std::string si = "SI";
            if (r.units.factor("DIST", f[3], si, fac)) r.units.vox *= fac;
            else st = -2;
        } else if (has(key, "ROUGH")) {
            if (n < 3) return too_few();
            if (!read_real(f[2], r.p.roughness)) return bad_field(3);
        } else if (has(key, "MAP")) {
            if (n < 4) return too_few();
            if (!read_real(f[3], r.p.avg_factor)) return bad_field(4);
        } else if (has(key, "CALCULATE")) {
            if (n < 2) return too_few();
            if (has(upper(f[1]), "PERCENT")) r.percentile = true;
            if (n >= 3) {
                r.perc.clear();
                for (int k = 2; k < n; k++) {
                    double v;
                    if (!read_real(f[k], v)) return bad_field(k + 1);
                    r.perc.push_back(v);
                }
            }
        } else if (has(key, "HIGH-MASK")) {
            if (n < 2) return too_few();
            if (!read_real(f[1], r.p.high_mask)) return bad_field(2);
        } else if (has(key, "LOW-MASK")) {
            if (n < 2) return too_few();
            if (!read_real(f[1], r.p.low_mask)) return bad_field(2);
        } else if (has(key, "MAXIMUM")) {
            if (n < 4) return too_few();
            const std::string t = f[3].substr(0, 4);   // "(I4)"
            char *end = nullptr;
            const long v = strtol(t.c_str(), &end, 10);
            if (end == t.c_str() || *end != '\0') return bad_field(4);
            r.maxdim = (int)v;
        } else if (has(key, "OUTPUT-UNITS")) {
            if (n < 4) return too_few();
            const char *types[3] = {"PRES", "DIST", "FLOW"};
            const int slots[3] = {1, 2, 5};
            for (int k = 0; k < 3; k++) {
                double fac;
                std::string si = "SI";
                r.unit_out[slots[k]] = f[k + 1];
                if (!r.units.factor(types[k], f[k + 1], si, fac)) return bad_field(k + 2);
            }
            output_unit = true;
        } else {
            r.message(" ** WARNING: UNRECOGNIZED INPUT: " + rtrim(rec.substr(0, 120)) + " **");
        }
        if (st == -1) return too_few();
        if (st > 0) return bad_field(st);
        if (st == -2) { /* CONVERT_VALUE set BOK false but READ_MASTER carries on */ r.p.flags |= 1; }
    }
    if (!output_unit)
        for (int k = 1; k <= 7; k++) r.unit_out[k] = r.unit_in[k];
    return true;
}

// One record of a data map under LNPROC's rules (no quotes in the file): fields end at blanks / '='
// (no empty fields) or ',' / TAB (two in a row make an empty field), ';' starts a comment.
// Returns the number of fields; *bad = 1-based index of a field that is not a number (0 = none).
int parse_map_line(const char *p, const char *end, int maxfld, std::vector<double> &vals, int *bad)
{
    vals.clear();
    *bad = 0;
    int state = 0;   // 1 inside a field, 0 after ',' / TAB, -1 after a blank or '='
    const char *fs = nullptr;
    int nfld = 0;
    auto close = [&](const char *fe) {
        // plain decimal fields ("12.5", "1e0", "-3.E-2", "+.5") go through std::from_chars -- correctly rounded like
        // strtod, a sixth of its cost, and a whole map is 1e7..1e8 of them; anything else (D exponents, stray letters,
        // out-of-range values, signs in odd places) takes the Fortran-minded path below
        {
            const char *q = fs;
            bool plain = fe > fs;
            for (const char *c = fs; c < fe && plain; c++)
                plain = (*c >= '0' && *c <= '9') || *c == '.' || *c == '-' || *c == '+' || *c == 'e' || *c == 'E';
            if (plain && *q == '+') q++;
            if (plain && q < fe && *q != '+') {
                double v;
                const auto res = std::from_chars(q, fe, v);
                if (res.ec == std::errc() && res.ptr == fe) { vals.push_back(v); return; }
            }
        }
        char tmp[128];
        size_t len = (size_t)(fe - fs);
        if (len >= sizeof tmp) len = sizeof tmp - 1;
        memcpy(tmp, fs, len);
        tmp[len] = '\0';
        double v;
        if (!read_real(tmp, v) && !*bad) *bad = nfld;
        vals.push_back(v);
    };
    for (; p <= end; p++) {
        const char c = p < end ? *p : ' ';
        if (c == ';') { if (state == 1) close(p); state = -1; break; }
        if (c == ' ' || c == '=' || c == '\r') {
            if (state == 1) { close(p); state = -1; }
        } else if (c == ',' || c == '\t') {
            if (state == 1) { close(p); state = 0; }
            else if (state == 0) {
                if (nfld == maxfld) return nfld;
                nfld++;
                vals.push_back(0.0);
                if (!*bad) *bad = nfld;          // a blank field cannot be read as a number
            } else state = 0;
        } else if (state != 1) {
            if (nfld == maxfld) return nfld;
            nfld++;
            fs = p;
            state = 1;
        }
    }
    return nfld;
}

// READ_DATA_MAP, MAP_MODULE.F:18-85.  The whole file is read once and its records are parsed by a pool
// of threads (records are independent when the file has no double quotes; otherwise the sequential
// LNPROC reader with its carried-over quote state is used).
bool read_data_map(Run &r)
{
    std::ifstream in(r.apm_file, std::ios::binary);
    if (!in) {
        r.to_stdout("IO ERROR: " + std::to_string(errno) + " ATTEMPTING TO OPEN: " + r.apm_file);
        return false;
    }
    std::string text;
    in.seekg(0, std::ios::end);
    const std::streamoff size = in.tellg();
    if (size > 0) {             // one read of the whole file (the iterator pair copies byte by byte: 0.3 GB/s)
        text.resize((size_t)size);
        in.seekg(0, std::ios::beg);
        in.read(&text[0], size);
        text.resize((size_t)in.gcount());
    } else {                    // not seekable: whatever the stream gives
        in.clear();
        in.seekg(0, std::ios::beg);
        text.assign((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    }
    in.close();
    std::vector<std::vector<double>> rows;
    int nx = 0;
    auto horizontal = [&]() { r.message(" ERROR - MAXIMUM NUMBER OF HORIZONTAL CELLS IS " + std::to_string(r.maxdim)); return false; };
    auto vertical = [&]() { r.message(" ERROR - MAXIMUM NUMBER OF VERTICAL CELLS IS " + std::to_string(r.maxdim)); return false; };
    if (text.find('"') != std::string::npos) {
        std::istringstream iss(text);
        LineReader lr(iss);
        std::vector<std::string> f;
        std::string rec;
        int n;
        while ((n = lr.next(f, rec, r.maxdim + 1)) > 0) {
            if (n > r.maxdim) return horizontal();
            if ((int)rows.size() + 1 > r.maxdim) return vertical();
            rows.emplace_back((size_t)n);
            for (int k = 0; k < n; k++)
                if (!read_real(f[k], rows.back()[k])) {
                    field_error(r, "APERTURE MAP", k + 1, "LINE - " + fixed(rec, 70) + "...");
                    return false;
                }
            nx = n;
        }
    } else {
        std::vector<std::pair<size_t, size_t>> lines;   // [begin, end) of every record
        for (size_t b = 0; b < text.size();) {
            size_t e = text.find('\n', b);
            if (e == std::string::npos) e = text.size();
            lines.emplace_back(b, e);
            b = e + 1;
        }
        const size_t nl = lines.size();
        std::vector<std::vector<double>> parsed(nl);
        std::vector<int> nf(nl, 0), bad(nl, 0);
        const unsigned T = nl > 256 ? io_threads() : 1;
        parallel_ranges((long long)nl, T, [&](long long b, long long e, unsigned) {
            for (long long k = b; k < e; k++)
                nf[(size_t)k] = parse_map_line(text.data() + lines[(size_t)k].first, text.data() + lines[(size_t)k].second,
                                               r.maxdim + 1, parsed[(size_t)k], &bad[(size_t)k]);
        });
        // replay the records in order so the first failure is the one the reference would report
        for (size_t k = 0; k < nl; k++) {
            if (nf[k] == 0) continue;   // blank / comment-only record
            if (nf[k] > r.maxdim) return horizontal();
            if ((int)rows.size() + 1 > r.maxdim) return vertical();
            if (bad[k]) {
                field_error(r, "APERTURE MAP", bad[k],
                            "LINE - " + fixed(upper(rtrim(text.substr(lines[k].first, lines[k].second - lines[k].first))), 70) + "...");
                return false;
            }
            rows.emplace_back(std::move(parsed[k]));
            nx = nf[k];   // the width of the LAST record wins, like the reference
        }
    }
    r.nz = (int)rows.size();
    r.nx = nx;
    if (r.nz == 0 || r.nx == 0) {
        r.message(" ERROR - APERTURE MAP IS EMPTY");
        return false;
    }
    r.ap.assign((size_t)r.nx * r.nz, 0.0);
    for (int z = 0; z < r.nz; z++)
        for (int x = 0; x < r.nx && x < (int)rows[z].size(); x++) r.ap[(size_t)z * r.nx + x] = rows[z][x];
    return true;
}

// CALC_PERCENTILE, OUTPUT_MODULE.F:53-124, and MAP_STATS, MAP_MODULE.F:182-303, run on the device (stats.cu: exact
// radix select instead of six sorts of N doubles, variogram lags in batches of eight); a failure there fails the
// run -- there is no host version to fall back to.
std::vector<double> percentiles(const std::vector<double> &v, const std::vector<double> &pct, int device)
{
    std::vector<double> res(pct.size());
    if (apm_lcl_percentiles(v.data(), v.size(), pct.data(), (int)pct.size(), device, res.data()))
        throw std::runtime_error(apm_lcl_last_error());
    return res;
}

void map_stats(const std::vector<double> &m, int nx, int nz, int device, double &rms, double &dev, double &xc, double &zc)
{
    if (apm_lcl_map_stats(m.data(), nx, nz, device, nullptr, &rms, &dev, &xc, &zc)) throw std::runtime_error(apm_lcl_last_error());
}

void write_percentile_table(Run &r, const std::vector<double> &pct, const std::vector<double> &val, const std::string &title,
                            const std::string &unit)
{
    const std::string rule = std::string(9, ' ') + std::string(32, '-');
    r.blank();
    r.message(rule);
    r.message(std::string(9, ' ') + title + " RANGES");
    r.message(rule);
    r.message(" " + A(18, "PERCENTILE") + "   |" + A(17, "[" + unit + "]"));
    r.message(rule);
    for (size_t k = 0; k < pct.size(); k++)
        r.message(std::string(9, ' ') + "|  " + F(7, 3, pct[k]) + "   | " + F(15, 6, val[k]) + " |");
    r.message(rule);
}

void write_map_csv(FILE *f, const double *v, int nx, int nz, double scale)
{
    write_rows_parallel(f, nz, [=](long long z, std::string &line) {
        for (int x = 0; x < nx; x++) {
            if (x) line += ',';
            append_G(line, 15, 8, v[(size_t)z * nx + x] * scale);
        }
        line += '\n';
    });
}

void write_vtk_scalars(Run &r, const char *name, const std::string &unit, const char *dtype, const double *v, double scale)
{
    FILE *f = r.fvtk;
    fprintf(f, " \nSCALARS %s[%s] %s\nLOOKUP_TABLE default\n", name, unit.c_str(), dtype);
    // "(*(1(E14.6,:,' '), /, :))" : one value per record, followed by the separator blank
    const int nx = r.nx;
    write_rows_parallel(f, r.nz, [=](long long z, std::string &buf) {
        for (int x = 0; x < nx; x++) {
            append_E(buf, 14, 6, v[(size_t)z * nx + x] * scale);
            buf += x + 1 < nx ? " \n" : "\n";
        }
    });
    r.message(std::string("         ") + name + " DATA OUTPUT TO VTK FILE");
}

double corner(const std::vector<double> &a, int nx, int nz, int cz, int cx)
{   // CREATE_POINT_MAP (MAP_MODULE.F:91-176) on the clamped map; corner whose lower-left cell is (cz,cx)
    auto at = [&](int z, int x) { return a[(size_t)z * nx + x]; };
    const bool zl = cz < 0, zh = cz >= nz - 1, xl = cx < 0, xh = cx >= nx - 1;
    if ((zl || zh) && (xl || xh)) return at(zl ? 0 : nz - 1, xl ? 0 : nx - 1);
    if (xl || xh) { const int x = xl ? 0 : nx - 1; return (at(cz, x) + at(cz + 1, x)) / 2.0; }
    if (zl || zh) { const int z = zl ? 0 : nz - 1; return (at(z, cx) + at(z, cx + 1)) / 2.0; }
    return (((at(cz, cx) + at(cz + 1, cx)) + at(cz, cx + 1)) + at(cz + 1, cx + 1)) / 4.0;
}

int side_code(const std::string &s)
{
    if (has(s, "LEFT")) return APM_LCL_LEFT;
    if (has(s, "RIGHT")) return APM_LCL_RIGHT;
    if (has(s, "TOP")) return APM_LCL_TOP;
    if (has(s, "BOTTOM")) return APM_LCL_BOTTOM;
    return -1;
}

int run(Run &r, const std::vector<int> &devices)
{
    // defaults, APERTURE_MAP_FLOW.F:88-126
    apm_lcl_default_params(&r.p);
    r.units.say = [&](const std::string &s) { r.to_stdout(s); };
    const char *si[8] = {"", "PA", "M", "K", "SEC", "M^3/SEC", "PA*SEC", "KG/M^3"};
    for (int k = 1; k <= 7; k++) { r.unit_in[k] = si[k]; r.unit_out[k] = si[k]; r.unit_si[k] = "SI"; }

    std::ifstream in;
    std::streampos resume = 0;
    if (!open_io(r, in, resume)) { r.open_io_failed = true; return 1; }
    in.clear();
    in.seekg(resume);
    if (!read_master(r, in)) return 1;
    in.close();

    const bool presc = r.p.inlet_pressure >= 0 && r.p.outlet_pressure >= 0;
    r.message(presc ? "     FRACTURE FLOW IS PRESSURE CONTROLLED" : "     FRACTURE FLOW IS RATE CONTROLLED");
    if (r.p.viscosity > 0) {
        r.message("     AVERAGE LIQUID VISCOSITY USED " + G(12, 6, r.p.viscosity / r.visc_conv) + " " + rtrim(r.unit_in[6]));
    } else {
        r.message(" ERROR - NONE OR INVAILD VISCOSITY PROVIDED " + G(15, 6, r.p.viscosity / r.visc_conv));
        return 1;
    }
    // input conversion factors (:156-161); UNIT(:) = 'SI' becomes the SI unit names
    bool ok = true;
    ok &= r.units.factor("PRES", r.unit_in[1], r.unit_si[1], r.pres_conv);
    ok &= r.units.factor("DIST", r.unit_in[2], r.unit_si[2], r.dim_conv);
    ok &= r.units.factor("TIME", r.unit_in[4], r.unit_si[4], r.time_conv);
    ok &= r.units.factor("RATE", r.unit_in[5], r.unit_si[5], r.rate_conv);
    ok &= r.units.factor("VISC", r.unit_in[6], r.unit_si[6], r.visc_conv);
    ok &= r.units.factor("DENS", r.unit_in[7], r.unit_si[7], r.rho_conv);
    if (!ok || (r.p.flags & 1)) return 1;
    r.p.flags = 0;

    // READ_AP_MAP :411-591
    r.blank();
    r.message(" READING APERTURE MAP");
    StageTimer timer;
    if (!read_data_map(r)) return 1;
    timer.lap("read + parse map");
    const int nx = r.nx, nz = r.nz;
    const size_t n = (size_t)nx * nz;
    const int dev0 = devices.empty() ? 0 : devices[0];
    r.p.voxel_size = r.units.vox;
    const double vox = r.units.vox;
    std::vector<double> perc_ap;
    if (r.percentile) {
        r.message("     CALCULATING APERTURE PERCENTILES OF MAP");
        perc_ap = percentiles(r.ap, r.perc, dev0);
        timer.lap("aperture percentiles");
    }
    r.p.outlet_side = side_code(r.outlet_side);

    // the hot path (device): everything numeric between READ_DATA_MAP and the writers
    std::vector<double> pres(n), qx(n), qz(n), qm(n), width(n), ren(n), res(n);
    apm_lcl_fields fo{pres.data(), qx.data(), qz.data(), qm.data(), width.data(), ren.data(), res.data()};
    apm_lcl_scalars sc{};
    int solve_rc = 0;
    std::string solve_err;
    if (r.p.outlet_side >= 0) {
        // the file pipeline prints 8-12 significant digits and is dominated by ASCII I/O: converge two
        // orders tighter than the tensor API's 1e-10 default so the printed digits are the solution's
        if (!(r.p.rel_tol > 0)) r.p.rel_tol = 1e-12;
        // several devices: the map is cut into row strips over those GPUs
        if (devices.size() > 1)
            solve_rc = apm_lcl_solve_striped(&r.p, r.ap.data(), nx, nz, (int)devices.size(), devices.data(), &fo, &sc);
        else solve_rc = apm_lcl_solve(&r.p, r.ap.data(), nx, nz, devices.empty() ? 0 : devices[0], &fo, &sc);
        if (solve_rc) solve_err = apm_lcl_last_error();
        timer.lap("apm_lcl_solve (host buffers)");
    }

    r.message("     CONVERTING 2-D MAP ARRAY IN VOX TO VECTOR IN SI");
    r.message("     ROUGHNESS FACTOR APPLIED TO MAP " + F(10, 6, r.p.roughness));
    r.message("     CALCULATING ADDITIONAL MAP STATISTICS");
    std::vector<double> clamped(n);
    for (size_t i = 0; i < n; i++) {
        double a = r.ap[i];
        if (a > r.p.high_mask) a = r.p.high_mask;
        if (a < r.p.low_mask) a = r.p.low_mask;
        clamped[i] = a;
    }
    {
        std::vector<double> si_map(n);
        for (size_t i = 0; i < n; i++) si_map[i] = clamped[i] * vox;
        double xc, zc;
        map_stats(si_map, nx, nz, dev0, r.aper_rms, r.aper_dev, xc, zc);
        timer.lap("map_stats");
        r.x_corr = xc * r.p.avg_factor * vox;
        r.z_corr = zc * r.p.avg_factor * vox;
    }
    double diam = nx * vox * r.p.avg_factor, leng = nz * vox * r.p.avg_factor;
    // scalars that do not need the device if the solve was refused (invalid side): recompute on host
    long long zeros = 0, masked = 0;
    double raw = 0, csum = 0, mn = 0, mx = r.ap[0];
    bool have_mn = false;
    for (int x = 0; x < nx; x++)
        for (int z = 0; z < nz; z++) {
            const double a = r.ap[(size_t)z * nx + x];
            raw += a;
            csum += clamped[(size_t)z * nx + x];
            if (std::fabs(a * 1e9) < 1.0) zeros++;
            if (a > mx) mx = a;
            if (a > 0 && (!have_mn || a < mn)) { mn = a; have_mn = true; }
            masked += (a > r.p.high_mask) + (a < r.p.low_mask);
        }
    raw /= (double)nx * nz;
    double avg_aper = csum / ((double)nz * nx) * vox;
    if (zeros > 0) {
        r.blank();
        r.message("     *** WARNING: " + std::to_string(zeros) + " CELLS HAD AN APERTURE OF ZERO ***");
        if (r.p.low_mask < 0.0001) r.message("     *** WARNING: A LOW MASK OF AT LEAST 0.0001 IS RECOMMENDED ***");
    }
    r.blank();
    r.message("     PROPERTIES OF APERTURE MAP:");
    r.blank();
    r.message(std::string(9, ' ') + "VOXEL TO METER CONVERSION USED: " + E(17, 6, vox));
    r.message(std::string(9, ' ') + "MAP AVERAGING FACTOR APPLIED: " + F(15, 6, r.p.avg_factor));
    r.blank();
    r.message("   " + A(17, "DIAMETER[M]") + " " + A(19, "LENGTH[M]") + std::string(9, ' ') + "NX" + std::string(8, ' ') + "NZ");
    r.message(std::string(5, ' ') + F(15, 6, diam) + std::string(5, ' ') + F(15, 6, leng) + std::string(5, ' ') + I(6, 5, nx) +
              std::string(4, ' ') + I(6, 5, nz));
    r.blank();
    r.message(std::string(9, ' ') + "NUMBER OF CELLS THAT WERE MASKED: " + std::to_string(masked));
    r.blank();
    r.message(std::string(9, ' ') + "MINIMUM, NONZERO APERTURE OF FRACTURE: " + F(15, 6, mn) + " VOXELS");
    r.message(std::string(9, ' ') + "MAXIMUM APERTURE OF FRACTURE: " + F(15, 6, mx) + " VOXELS");
    r.blank();
    r.message(std::string(9, ' ') + "HOMOGENEOUS APERTURE OF FRACTURE " + A(22, "WITHOUT MASK APPLIED: ") + F(15, 6, raw) + " VOXELS");
    r.message(std::string(9, ' ') + "HOMOGENEOUS APERTURE OF FRACTURE " + A(22, "WITH MASK APPLIED:    ") + F(15, 6, avg_aper / vox) + " VOXELS");
    if (r.percentile) {
        r.blank();
        r.message(std::string(9, ' ') + "APERTURE RANGES BY PERCENTILE");
        r.message(" " + A(18, "PERCENTILE") + " " + A(19, "[VOX]") + " " + A(19, "[MM]"));
        for (size_t k = 0; k < r.perc.size(); k++)
            r.message(std::string(12, ' ') + F(7, 3, r.perc[k]) + std::string(5, ' ') + F(15, 6, perc_ap[k]) + std::string(5, ' ') +
                      F(15, 9, perc_ap[k] * vox * 1000));
    }
    // FLOW_BOUNDARY :277-376
    r.blank();
    r.message(" DEFINING BOUNDARY CELL TRANSMISSIBILIES");
    if (r.p.outlet_side < 0) {
        r.message(" INVALID OUTLET SIDE ENTERED: " + rtrim(r.outlet_side));
        return 1;
    }
    r.blank();
    r.message(" CALCULATING FLOW THROUGH FRACTURE");
    r.blank();
    {   // one line where the reference prints "USING D4 GAUSS SOLVER", so the log keeps its line count
        char buf[200];
        snprintf(buf, sizeof buf, "     USING B200 MULTIGRID-PCG SOLVER (%d ITERATIONS, RESIDUAL %s)", sc.iterations,
                 ffmt::rtrim(E(12, 4, sc.rel_residual)).c_str());
        r.message(solve_rc ? "     USING B200 MULTIGRID-PCG SOLVER" : buf);
    }
    if (solve_rc) {
        r.message(" ERROR - " + solve_err);
        return 1;
    }
    {
        char buf[200];
        snprintf(buf, sizeof buf, "     REQUIRED TIME FOR SOLUTION: %s", F(13, 6, sc.total_ms * 1e-3).c_str());
        r.message(buf);
    }
    r.blank();
    r.message("      FRACTURE FLOWS CALCULATED");

    // INIT_OUTPUT :945-1040
    double conv_in[8] = {0, r.pres_conv, r.dim_conv, 0, r.time_conv, r.rate_conv, r.visc_conv, r.rho_conv};
    r.unit_out[6] = r.unit_in[6];
    r.unit_out[7] = r.unit_in[7];
    r.blank();
    r.message(" INITIALIZING OUTPUT FILES");
    r.message("     INPUT CONVERSIONS APPLIED");
    ok = true;
    ok &= r.units.factor("PRES", r.unit_si[1], r.unit_out[1], r.pres_conv);
    ok &= r.units.factor("DIST", r.unit_si[2], r.unit_out[2], r.dim_conv);
    ok &= r.units.factor("TIME", r.unit_si[4], r.unit_out[4], r.time_conv);
    ok &= r.units.factor("RATE", r.unit_si[5], r.unit_out[5], r.rate_conv);
    ok &= r.units.factor("VISC", r.unit_si[6], r.unit_out[6], r.visc_conv);
    ok &= r.units.factor("DENS", r.unit_si[7], r.unit_out[7], r.rho_conv);
    if (!ok) return 1;
    const std::string rule = std::string(5, ' ') + std::string(41, '-');
    auto table = [&](const std::string *from, const std::string *to, const double *cv) {
        r.message(rule);
        r.message(std::string(5, ' ') + A(10, "UNIT IN") + std::string(4, ' ') + A(10, "UNIT OUT") + "  " + A(15, "CONVERSION"));
        r.message(rule);
        for (int k = 1; k <= 7; k++) {
            if (k == 3) continue;
            r.message(std::string(5, ' ') + fixed(from[k], 10) + " -> " + fixed(to[k], 10) + ": " + G(15, 6, cv[k]));
        }
        r.message(rule);
    };
    table(r.unit_in, r.unit_si, conv_in);
    double conv_out[8] = {0, r.pres_conv, r.dim_conv, 0, r.time_conv, r.rate_conv, r.visc_conv, r.rho_conv};
    r.blank();
    r.message("     OUTPUT CONVERSIONS APPLIED");
    table(r.unit_si, r.unit_out, conv_out);

    // OUTPUT_DATA :861-896
    double inl = sc.inlet_pressure, outp = sc.outlet_pressure;
    double outrate = sc.outlet_rate * r.rate_conv;
    r.blank();
    r.message("     INLET BOUNDARY PRESSURE: " + G(20, 9, inl * r.pres_conv) + " " + rtrim(r.unit_out[1]));
    r.message("     OUTLET BOUNDARY PRESSURE: " + G(20, 9, outp * r.pres_conv) + " " + rtrim(r.unit_out[1]));
    r.message("     NET FLOW RATE CALCULATED: " + G(20, 9, outrate) + " " + rtrim(r.unit_out[5]));
    r.message("     MAXIMUM NET FLOW RESIDUAL: " + G(20, 9, sc.max_residual));
    if (r.percentile) {
        // an order statistic of the scaled field is the scaled order statistic (the unit factor is positive and
        // rounding is monotone): select on the SI field, convert the handful of results
        auto scaled = [&](std::vector<double> v, double s) { for (auto &e : v) e *= s; return v; };
        write_percentile_table(r, r.perc, scaled(percentiles(qx, r.perc, dev0), r.rate_conv), "X-FLOW", rtrim(r.unit_out[5]));
        write_percentile_table(r, r.perc, scaled(percentiles(qz, r.perc, dev0), r.rate_conv), "Z-FLOW", rtrim(r.unit_out[5]));
        write_percentile_table(r, r.perc, scaled(percentiles(qm, r.perc, dev0), r.rate_conv), "FLOW MAGNITUDE", rtrim(r.unit_out[5]));
        write_percentile_table(r, r.perc, percentiles(res, r.perc, dev0), "NORMALIZED RESIDUAL", "-");
    }
    r.blank();
    timer.lap("scalars, log, percentiles");
    r.message(" WRITING DATA TO OUTPUT FILES...");

    // WRITE_STATS_FILE :1046-1209
    {
        const double visc = r.p.viscosity, rho = r.p.density, flow = sc.flow;
        const double avg_trx = (leng * avg_aper * avg_aper * avg_aper) / (12.0 * diam) / visc;
        const double avg_trz = (diam * avg_aper * avg_aper * avg_aper) / (12.0 * leng) / visc;
        const bool horiz = has(r.outlet_side, "RIGHT") || has(r.outlet_side, "LEFT");
        double flow_cf, nfq, hom, hyd, bulk;
        if (horiz) {
            flow_cf = std::fabs(flow / (inl - outp) / avg_trx);
            nfq = std::fabs(avg_trx * (inl - outp));
            hom = std::fabs(flow / (inl - outp)) * visc * diam;
            hyd = std::fabs(flow / (inl - outp)) * 12.0 * visc * diam / leng;
            bulk = std::fabs(2.0 * rho * flow) / (visc * leng);
        } else {
            flow_cf = std::fabs(flow / (inl - outp) / avg_trz);
            nfq = std::fabs(avg_trz * (inl - outp));
            hom = std::fabs(flow / (inl - outp)) * visc * leng;
            hyd = std::fabs(flow / (inl - outp)) * 12.0 * visc * leng / diam;
            bulk = std::fabs(2.0 * rho * flow) / (visc * diam);
        }
        hyd = std::pow(hyd, 1.0 / 3.0);
        const double dc = r.dim_conv;
        const double leng_o = leng * dc, diam_o = diam * dc, avg_o = avg_aper * dc, hyd_o = hyd * dc;
        const double visc_o = visc * r.visc_conv, rho_o = rho * r.rho_conv;
        const double rms_o = r.aper_rms * dc, dev_o = r.aper_dev * dc, zc_o = r.z_corr * dc, xc_o = r.x_corr * dc;
        const double inl_o = inl * r.pres_conv, out_o = outp * r.pres_conv, nfq_o = nfq * r.rate_conv;
        const double avg_corr = (2.0 * xc_o * zc_o) / (xc_o + zc_o);
        const double dp = (inl_o - out_o) / r.pres_conv;
        const std::string u1 = rtrim(r.unit_out[1]), u2 = rtrim(r.unit_out[2]), u5 = rtrim(r.unit_out[5]),
                          u6 = rtrim(r.unit_out[6]), u7 = rtrim(r.unit_out[7]);
        auto row = [&](std::initializer_list<double> v) {
            std::string s;
            bool first = true;
            for (double x : v) { if (!first) s += ','; s += G(20, 12, x); first = false; }
            return s + "\n";
        };
        FILE *c = r.fstatc;
        fprintf(c, "APER-MAP:, %s\nNX,NZ\n%d,%d\n", r.apm_file.c_str(), nx, nz);
        fprintf(c, "LENGTH [%s],DIAMETER [%s],MEAN APER [%s],\n%s", u2.c_str(), u2.c_str(), u2.c_str(), row({leng_o, diam_o, avg_o}).c_str());
        fprintf(c, "AVG VISC [%s],DENSITY [%s]\n%s", u6.c_str(), u7.c_str(), row({visc_o, rho_o}).c_str());
        fprintf(c, "INLET PRESS [%s],OUTLET PRESS [%s],OUTLET RATE [%s]\n%s", u1.c_str(), u1.c_str(), u5.c_str(), row({inl_o, out_o, outrate}).c_str());
        fprintf(c, "#\nAPERTURE RMS [%s],APERTURE STDDEV [%s]\n%s", u2.c_str(), u2.c_str(), row({rms_o, dev_o}).c_str());
        fprintf(c, "#\nX-CORR LENGTH [%s],Z-CORR LENGTH [%s],AVG-CORR LENGTH [%s]\n%s", u2.c_str(), u2.c_str(), u2.c_str(), row({xc_o, zc_o, avg_corr}).c_str());
        fprintf(c, "#\nSTDDEV/MEAN APER [-],STDDEV/AVG CORR LEN [-], MEAN APER/AVG CORR LEN [-]\n%s", row({dev_o / avg_o, dev_o / avg_corr, avg_o / avg_corr}).c_str());
        fprintf(c, "#\nBULK REYNOLDS # [-],AVERAGE TRANSMISSIBILTY [M^4],PRESSURE DIFFERENTIAL [PA]\n%s", row({bulk, hom, dp}).c_str());
        fprintf(c, "#\n1-D EQUIV [%s],PERC OF 1-D [-]\n%s", u5.c_str(), row({nfq_o, flow_cf * 100}).c_str());
        FILE *y = r.fstaty;
        auto yl = [&](const char *key, double v, const std::string &u) { fprintf(y, "%s: [%s,\"%s\"]\n", key, G(20, 12, v).c_str(), u.c_str()); };
        fprintf(y, "APER-MAP: %s\nNX: %d\nNZ: %d\n", r.apm_file.c_str(), nx, nz);
        yl("LENGTH", leng_o, u2); yl("DIAMETER", diam_o, u2); yl("MEAN APER", avg_o, u2); yl("HYDRAULIC APER", hyd_o, u2);
        yl("AVG VISC", visc_o, u6); yl("DENSITY", rho_o, u7); yl("INLET PRESS", inl_o, u1); yl("OUTLET PRESS", out_o, u1);
        yl("OUTLET RATE", outrate, u5); yl("APERTURE RMS", rms_o, u2); yl("APERTURE STDDEV", dev_o, u2);
        yl("X-CORR LENGTH", xc_o, u2); yl("Z-CORR LENGTH", zc_o, u2); yl("AVG-CORR LENGTH", avg_corr, u2);
        yl("STDDEV/MEAN APER", dev_o / avg_o, "-"); yl("STDDEV/AVG CORR LEN", dev_o / avg_corr, "-");
        yl("MEAN APER/AVG CORR LEN", avg_o / avg_corr, "-"); yl("BULK REYNOLDS NUM", bulk, "-");
        yl("AVERAGE TRANSMISSIBILTY", hom, "M^4"); yl("PRESSURE DIFFERENTIAL", dp, "PA"); yl("1-D EQUIV", nfq_o, u5);
        yl("PERC OF 1-D", flow_cf * 100, "-");
    }
    timer.lap("stat files");
    write_map_csv(r.fmap, width.data(), nx, nz, r.dim_conv);
    r.message("     ADJUSTED APERTURE FILE COMPLETED");
    write_map_csv(r.fflox, qx.data(), nx, nz, r.rate_conv);
    write_map_csv(r.ffloz, qz.data(), nx, nz, r.rate_conv);
    write_map_csv(r.fflom, qm.data(), nx, nz, r.rate_conv);
    r.message("     FLOW FILE COMPLETED");
    write_map_csv(r.fpres, pres.data(), nx, nz, r.pres_conv);
    r.message("     PRESSURE FILE COMPLETED");
    timer.lap("5 csv maps");

    // WRITE_VTK_FILE :1215-1291
    {
        FILE *f = r.fvtk;
        r.blank();
        r.message("     CREATING THE VTK FILE...");
        fprintf(f, "# vtk DataFile Version 3.0\nvtk output\nASCII\nDATASET STRUCTURED_GRID\nDIMENSIONS %d %d 2\nPOINTS %lld float\n",
                nx + 1, nz + 1, (long long)(nx + 1) * (nz + 1) * 2);
        const double pm = vox * r.dim_conv;
        const double dx = diam / nx * r.dim_conv, dz = leng / nz * r.dim_conv;
        // X and Z are running sums in the reference (X = X + DX); reproduce the same partial sums
        std::vector<double> xs((size_t)nx + 1), zs((size_t)nz + 1);
        xs[0] = 0.0;
        for (int ix = 0; ix < nx; ix++) xs[ix + 1] = xs[ix] + dx;
        zs[0] = 0.0;
        for (int iz = 0; iz < nz; iz++) zs[iz + 1] = zs[iz] + dz;
        // the x coordinate of a column and the z coordinate of a row are the same text in every record that shows them:
        // formatted once (two thirds of the conversions of this section)
        std::string xtext;
        for (int ix = 0; ix <= nx; ix++) { append_E(xtext, 14, 6, xs[(size_t)ix]); xtext += ' '; }
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            write_rows_parallel(f, (long long)nz + 1, [&, sgn](long long iz, std::string &buf) {
                // rows 0..nz-1 use the bottom corners of row iz, the last record the top corners of row nz-1
                const int cz = iz < nz ? (int)iz - 1 : nz - 1;
                std::string ztext(1, ' ');
                append_E(ztext, 14, 6, zs[(size_t)iz]);
                ztext += '\n';
                for (int ix = -1; ix < nx; ix++) {
                    buf.append(xtext, (size_t)(ix + 1) * 15, 15);
                    append_E(buf, 14, 6, sgn * (corner(clamped, nx, nz, cz, ix) * pm) / 2.0);
                    buf += ztext;
                }
            });
        }
        r.message("         VTK POINTS OUTPUT TO FILE");
        r.blank();
        r.message("     OUTPUTTING DATA TO THE VTK FILE...");
        fprintf(f, " \nCELL_DATA %lld\n", (long long)nx * nz);
        write_vtk_scalars(r, "APERTURE", rtrim(r.unit_out[2]), "float", width.data(), r.dim_conv);
        write_vtk_scalars(r, "PRESSURE", rtrim(r.unit_out[1]), " float", pres.data(), r.pres_conv);
        write_vtk_scalars(r, "REYNOLDS_NUMBER", "-", "float", ren.data(), 1.0);
        write_vtk_scalars(r, "NORMALIZED_RESIDUAL", "-", "float", res.data(), 1.0);
        fprintf(f, " \nVECTORS FLOW[%s] float\n", rtrim(r.unit_out[5]).c_str());
        {
            const double rc = r.rate_conv;
            const double *px = qx.data(), *pz = qz.data();
            write_rows_parallel(f, nz, [=](long long z, std::string &b2) {
                for (int x = 0; x < nx; x++) {
                    const size_t i = (size_t)z * nx + x;
                    append_E(b2, 14, 6, px[i] * rc);
                    b2 += "   0.000000E+00 ";
                    append_E(b2, 14, 6, pz[i] * rc);
                    b2 += '\n';
                }
            });
        }
        r.message("         VECTOR FLOW DATA OUTPUT TO VTK FILE");
    }
    timer.lap("vtk file");
    r.blank();
    r.message(" DEALLOCATED ALL ARRAYS");
    r.blank();
    r.message(" OUTPUT FILES CLOSED");
    return 0;
}

}   // namespace

static int run_file(const char *inp_path, const std::vector<int> &devices, char *log, size_t log_cap)
{
    Run r;
    int rc = 1;
    if (!inp_path || !*inp_path) {
        r.to_stdout("FATAL ERROR - NO INPUT FILE SUPPLIED");   // IO_MODULE.F:128
    } else {
        r.inp = inp_path;
        try {
            rc = run(r, devices);
        } catch (const std::exception &e) {
            r.to_stdout(std::string(" ERROR - ") + e.what());
            rc = 1;
        }
    }
    // a failed OPEN_IO removes what it created (IO_MODULE.F:219-226); later failures keep the files
    r.close_all(r.open_io_failed);
    r.to_stdout(" ");
    r.to_stdout("     **** EXITING PROGRAM ****");
    if (log && log_cap) {
        const size_t m = std::min(log_cap - 1, r.out.size());
        // keep the END of the transcript when it does not fit (error messages are at the end)
        memcpy(log, r.out.data() + (r.out.size() - m), m);
        log[m] = '\0';
    }
    return rc;
}

extern "C" APM_LCL_API int apm_lcl_run_file(const char *inp_path, int device, char *log, size_t log_cap)
{
    // APM_LCL_DEVICES="0,1,2,3" (the executable's way to name several GPUs) overrides `device`
    std::vector<int> devices;
    if (const char *e = getenv("APM_LCL_DEVICES")) {
        std::stringstream ss(e);
        std::string tok;
        while (std::getline(ss, tok, ','))
            if (!tok.empty()) devices.push_back(atoi(tok.c_str()));
    }
    if (devices.empty()) devices.push_back(device);
    return run_file(inp_path, devices, log, log_cap);
}

extern "C" APM_LCL_API int apm_lcl_run_file_striped(const char *inp_path, const int *devices, int n_devices, char *log,
                                                    size_t log_cap)
{
    std::vector<int> devs(devices, devices + (devices && n_devices > 0 ? n_devices : 0));
    if (devs.empty()) devs.push_back(0);
    return run_file(inp_path, devs, log, log_cap);
}
