// smoke_run.cpp -- headless scene runner on the C ABI.
//
// The reference's only runner is the interactive main loop (project/main.cpp:861-896): it steps when 1/20 s of WALL-CLOCK time
// has passed and passes that time as dt (main.cpp:891-895), so no two runs are alike.  This runner replays a scene with
// FIXED ticks -- first tick dt0 = 0.01 (main.cpp:293), then dt = 0.05 (main.cpp:93) unless the scene says otherwise -- and can
// dump every field as raw little-endian arrays in the reference's layout (x fastest; staggered fields (W+1)(H+1)(D+1),
// cu:146-147) for field-by-field comparison with any other implementation.  Plain C++ against include/smoke_b200.h:
//
//   g++ -O2 smoke_run.cpp -I../../include -L.. -lsmoke_b200 -Wl,-rpath,'$ORIGIN/..' -o smoke_run
//   ./smoke_run --scene C1 --ticks 20 --dump out/          named scenes C1..C4 and C5:<G> (scenes.py)
//   ./smoke_run --scene my.scene --dump out/ --dump-every 5
//
// Scene file (text, '#' comments):   grid W H D | gravity g | buoyancy a | iterations n | ticks n | dt0 s | dt s |
//   source x y z r | obstacle x y z r | union 0/1 | move <object id> <tick> x y z     (ids: creation order, sources and
//   obstacles share one sequence like cu:58, 91-94; "move" = updateObjectPos before that tick, main.cpp:609-618)
// Dump directory: density.f32  u.f32  v.f32  w.f32  mask.u8  meta.json   (with --dump-every k: tick_<t>/ subdirectories)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

#include "smoke_b200.h"

namespace {

struct Obj { int type; float x, y, z, r; };
struct Move { int id, tick; float x, y, z; };
struct Scene {
    unsigned W = 80, H = 80, D = 80;
    float gravity = -9.82f, buoyancy = 2.0f, dt0 = 0.01f, dt = 0.05f;
    int iterations = 30, ticks = 20, obstacle_union = 0;
    std::vector<Obj> objs;
    std::vector<Move> moves;
};

[[noreturn]] void die(const std::string& m)
{
    fprintf(stderr, "smoke_run: %s\n", m.c_str());
    exit(2);
}

// the deterministic scenes of SURVEY 8(d) (the same numbers as smoke_simulation_b200/scenes.py)
bool named_scene(const std::string& name, Scene& s)
{
    auto set = [&](unsigned W, unsigned H, unsigned D, float g, float a) { s.W = W; s.H = H; s.D = D; s.gravity = g; s.buoyancy = a; };
    if (name == "C1") { set(80, 80, 80, -9.82f, 2.0f); s.objs = {{1, 40, 40, 40, 5}, {0, 60, 10, 60, 13}}; return true; } // main.cpp:87, 288-291
    if (name == "C2") { set(256, 256, 256, -9.82f, 15.0f); s.objs = {{1, 128, 32, 128, 16}}; return true; }
    if (name == "C3") { set(512, 512, 512, -9.82f, 15.0f); s.objs = {{1, 256, 64, 256, 32}, {0, 256, 192, 256, 48}}; return true; }
    if (name == "C4") { set(1024, 1024, 1024, 9.82f, 2.0f); s.objs = {{1, 512, 128, 512, 64}}; return true; }
    if (name.rfind("C5", 0) == 0) {
        const unsigned G = name.size() > 3 ? (unsigned)atoi(name.c_str() + 3) : 1u;
        if (G < 1 || G > 64) return false;
        set(512, 512, 512 * G, -9.82f, 15.0f); s.objs = {{1, 256, 64, 256.0f * G, 32}};
        return true;
    }
    return false;
}

void parse_scene_file(const std::string& path, Scene& s)
{
    std::ifstream in(path);
    if (!in) die("cannot open scene file " + path);
    std::string line;
    int ln = 0;
    while (std::getline(in, line)) {
        ln++;
        const size_t h = line.find('#');
        if (h != std::string::npos) line.resize(h);
        std::istringstream is(line);
        std::string k;
        if (!(is >> k)) continue;
        bool ok = true;
        if (k == "grid") ok = (bool)(is >> s.W >> s.H >> s.D);
        else if (k == "gravity") ok = (bool)(is >> s.gravity);
        else if (k == "buoyancy") ok = (bool)(is >> s.buoyancy);
        else if (k == "iterations") ok = (bool)(is >> s.iterations);
        else if (k == "ticks") ok = (bool)(is >> s.ticks);
        else if (k == "dt0") ok = (bool)(is >> s.dt0);
        else if (k == "dt") ok = (bool)(is >> s.dt);
        else if (k == "union") ok = (bool)(is >> s.obstacle_union);
        else if (k == "source" || k == "obstacle") { Obj o{k == "source" ? 1 : 0, 0, 0, 0, 0}; ok = (bool)(is >> o.x >> o.y >> o.z >> o.r); s.objs.push_back(o); }
        else if (k == "move") { Move m{}; ok = (bool)(is >> m.id >> m.tick >> m.x >> m.y >> m.z); s.moves.push_back(m); }
        else ok = false;
        if (!ok) die(path + ":" + std::to_string(ln) + ": cannot parse '" + line + "'");
    }
}

void write_raw(const std::string& path, const void* p, size_t bytes)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f || fwrite(p, 1, bytes, f) != bytes) die("cannot write " + path);
    fclose(f);
}

void ck(int rc, smk_sim* s, const char* what)
{
    if (rc != SMK_OK) die(std::string(what) + ": " + smk_last_error(s));
}

void dump(smk_sim* sim, const Scene& sc, const std::string& dir, int tick, const std::vector<float>& density)
{
    mkdir(dir.c_str(), 0777);
    const size_t nc = (size_t)sc.W * sc.H * sc.D, ns = (size_t)(sc.W + 1) * (sc.H + 1) * (sc.D + 1);
    write_raw(dir + "/density.f32", density.data(), nc * 4);
    std::vector<float> f(ns);
    const char* names[3] = {"u", "v", "w"};
    for (int i = 0; i < 3; i++) {
        ck(smk_get_field(sim, SMK_FIELD_U + i, SMK_BUF_NOW, f.data()), sim, "smk_get_field");
        write_raw(dir + "/" + names[i] + ".f32", f.data(), ns * 4);
    }
    std::vector<unsigned char> m(nc);
    ck(smk_get_field(sim, SMK_FIELD_MASK, SMK_BUF_NOW, m.data()), sim, "smk_get_field(mask)");
    write_raw(dir + "/mask.u8", m.data(), nc);
    double sum = 0, mx = 0;
    for (float v : density) { sum += v; mx = std::fmax(mx, v); }
    float res = 0.f;
    ck(smk_max_divergence(sim, &res), sim, "smk_max_divergence");
    FILE* j = fopen((dir + "/meta.json").c_str(), "w");
    if (!j) die("cannot write meta.json");
    fprintf(j, "{\"grid\": [%u, %u, %u], \"ticks_done\": %d, \"density_sum\": %.9g, \"density_max\": %.9g, \"max_abs_divergence\": %.9g, "
               "\"layout\": \"x fastest; density/mask W*H*D, u/v/w (W+1)(H+1)(D+1) post-projection ('now') velocities, little-endian\"}\n",
            sc.W, sc.H, sc.D, tick, sum, mx, (double)res);
    fclose(j);
}

} // namespace

int main(int argc, char** argv)
{
    Scene sc;
    std::string scene = "C1", dump_dir;
    int ticks = -1, dump_every = 0;
    bool readback = true, quiet = false;
    for (int i = 1; i < argc; i++) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) die("missing value after " + a); return argv[++i]; };
        if (a == "--scene") scene = next();
        else if (a == "--ticks") ticks = atoi(next().c_str());
        else if (a == "--dump") dump_dir = next();
        else if (a == "--dump-every") dump_every = atoi(next().c_str());
        else if (a == "--no-readback") readback = false;   // simulate(nullptr, dt): the density stays on the device
        else if (a == "--quiet") quiet = true;
        else die("unknown option " + a + " (see the header of smoke_run.cpp)");
    }
    if (!named_scene(scene, sc)) parse_scene_file(scene, sc);
    if (ticks >= 0) sc.ticks = ticks;

    smk_sim* sim = nullptr;
    std::vector<float> density((size_t)sc.W * sc.H * sc.D, 0.f);
    if (smk_create(&sim, sc.W, sc.H, sc.D, density.data()) != SMK_OK) die(std::string("smk_create: ") + smk_last_error(nullptr));
    *smk_gravity_ptr(sim) = sc.gravity;
    *smk_buoyancy_ptr(sim) = sc.buoyancy;
    ck(smk_set_solver(sim, SMK_SOLVER_RBGS, sc.iterations, 0), sim, "smk_set_solver");
    ck(smk_set_obstacle_mode(sim, sc.obstacle_union), sim, "smk_set_obstacle_mode");
    for (const Obj& o : sc.objs) {
        const int id = o.type ? smk_add_source(sim, o.x, o.y, o.z, o.r) : smk_add_obstacle(sim, o.x, o.y, o.z, 0, 0, 0, o.r);
        if (id < 0) die(std::string("adding a scene object: ") + smk_last_error(sim));
    }
    for (int t = 0; t < sc.ticks; t++) {
        for (const Move& m : sc.moves)
            if (m.tick == t) ck(smk_update_object_pos(sim, m.id, m.x, m.y, m.z), sim, "smk_update_object_pos");
        const bool want = readback || (!dump_dir.empty() && (t == sc.ticks - 1 || (dump_every > 0 && (t + 1) % dump_every == 0)));
        ck(smk_step(sim, t == 0 ? sc.dt0 : sc.dt, want ? density.data() : nullptr), sim, "smk_step");
        if (!dump_dir.empty() && dump_every > 0 && (t + 1) % dump_every == 0 && t != sc.ticks - 1) {
            mkdir(dump_dir.c_str(), 0777);
            dump(sim, sc, dump_dir + "/tick_" + std::to_string(t + 1), t + 1, density);
        }
    }
    if (!dump_dir.empty()) dump(sim, sc, dump_dir, sc.ticks, density);
    double sum = 0, mx = 0;
    for (float v : density) { sum += v; mx = std::fmax(mx, v); }
    float res = 0.f;
    ck(smk_max_divergence(sim, &res), sim, "smk_max_divergence");
    if (!quiet)
        printf("smoke_run: scene %s, %ux%ux%u, %d ticks, sum(density) = %.6f, max = %.6f, max|div| = %.3e, %ld kernel launches\n", scene.c_str(),
               sc.W, sc.H, sc.D, sc.ticks, sum, mx, (double)res, smk_launch_count(sim));
    smk_destroy(sim);
    return 0;
}
