// host_demo.cpp — headless C++ host over csrc/world.hpp + libfse_b200.so: the reference's game loop order (game::tick,
// game.cpp:1659-2201) driven from C++ the way a maintainer's `class world` would drive it.
//
//   host_demo <table.bin> <world.bin> <W> <H> <ticks> [n_entities]
//
// table.bin / world.bin are raw dumps of the flattened material table (the argument list of fse_materials_set) and of the W x H
// fse_cell grid; tests/test_cpp_host.py writes them, runs this program on the GPU box and compares the printed state hash with the
// same loop driven through the Python binding.  No CUDA headers: the C ABI is all this file sees.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "world.hpp"

template <class T>
static void read_vec(std::ifstream& f, std::vector<T>& v) {
    int64_t n = 0;
    f.read(reinterpret_cast<char*>(&n), sizeof n);
    v.resize((size_t)n);
    if (n) f.read(reinterpret_cast<char*>(v.data()), sizeof(T) * (size_t)n);
}

int main(int argc, char** argv) {
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s table.bin world.bin W H ticks [n_entities]\n", argv[0]);
        return 2;
    }
    try {
        const int W = std::atoi(argv[3]), H = std::atoi(argv[4]), ticks = std::atoi(argv[5]), n_ent = argc > 6 ? std::atoi(argv[6]) : 0;
        fse_host::MaterialTable tbl;
        {
            std::ifstream f(argv[1], std::ios::binary);
            if (!f) throw fse_host::Error("cannot open the material table");
            read_vec(f, tbl.mats);
            f.read(reinterpret_cast<char*>(&tbl.ids), sizeof tbl.ids);
            read_vec(f, tbl.inter);
            read_vec(f, tbl.inter_offsets);
            read_vec(f, tbl.react);
            read_vec(f, tbl.react_offsets);
        }
        std::vector<fse_cell> cells((size_t)W * H);
        {
            std::ifstream f(argv[2], std::ios::binary);
            if (!f) throw fse_host::Error("cannot open the world dump");
            f.read(reinterpret_cast<char*>(cells.data()), sizeof(fse_cell) * cells.size());
        }
        fse_host::Context ctx(0);
        ctx.materials_push(tbl);
        fse_host::world w;
        w.init(ctx, W, H);
        fse_host::check(fse_write_rect(w.handle(), 0, 0, W, H, cells.data()));
        std::vector<fse_entity> ents;
        for (int i = 0; i < n_ent; i++) {
            fse_entity e{};
            e.x = 150.0f + 37.0f * i;
            e.y = 150.0f + 11.0f * i;
            e.vx = (i % 2) ? 1.5f : -1.0f;
            e.vy = 0.0f;
            e.hw = 8 + i;
            e.hh = 14 + 2 * i;
            ents.push_back(e);
        }
        std::vector<fse_xform> no_bodies;
        fse_render_stats moving{};
        for (int t = 0; t < ticks; t++) w.gameTick(no_bodies, ents, [](size_t, float, float) {}, &moving);
        fse_stats st{};
        w.stats(&st);
        int64_t np = 0;
        fse_host::check(fse_particles_count(w.handle(), &np));
        std::printf("hash=%016llx particles=%lld dirty_last_tick=%lld cut_outs=%zu", (unsigned long long)st.hash, (long long)np, (long long)moving.dirty,
                    w.cutOuts.size());
        for (const fse_entity& e : ents) std::printf(" ent=%.6f,%.6f,%.6f,%.6f,%d", e.x, e.y, e.vx, e.vy, e.ground);
        std::printf("\n");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "host_demo: %s\n", e.what());
        return 1;
    }
    return 0;
}
