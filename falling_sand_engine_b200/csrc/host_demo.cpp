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
        // fracture hand-off from C++: a cracked plate goes through updateRigidBodyHitbox (device pieces + outlines, host triangles)
        size_t hb_pieces = 0, hb_tris = 0;
        double hb_sum = 0;
        {
            const int bw = 40, bh = 24;
            std::vector<fse_cell> plate((size_t)bw * bh);
            for (int y = 0; y < bh; y++)
                for (int x = 0; x < bw; x++) {
                    fse_cell c{};
                    const bool solid = x != 17 && !(y >= 3 && y < 6 && x >= 25 && x < 28) && !((x * 7 + y * 13) % 11 == 0 && x > 30);
                    c.mat = solid ? 22 : 0;
                    c.color = 0x404040u + (uint32_t)(x + y * bw);
                    c.fluid = 2.0f;
                    plate[(size_t)x + (size_t)y * bw] = c;
                }
            fse_body_desc d{bw, bh, plate.data()};
            fse_host::check(fse_bodies_upload(w.handle(), &d, 1));
            for (const auto& pc : w.updateRigidBodyHitbox(0, bw, bh, 0.3f, 5, 5)) {
                hb_pieces++;
                for (const auto& grp : pc.shapes)
                    for (const auto& t : grp) {
                        hb_tris++;
                        for (int k = 0; k < 3; k++) hb_sum += t.p[k].x * (k + 1) + t.p[k].y * (k + 4) + pc.rec.x0 + 2.0 * pc.rec.y0;
                    }
            }
        }
        fse_stats st{};
        w.stats(&st);
        int64_t np = 0;
        fse_host::check(fse_particles_count(w.handle(), &np));
        std::printf("hash=%016llx particles=%lld dirty_last_tick=%lld cut_outs=%zu", (unsigned long long)st.hash, (long long)np, (long long)moving.dirty,
                    w.cutOuts.size());
        std::printf(" hitbox=%zu,%zu,%.3f", hb_pieces, hb_tris, hb_sum);
        for (const fse_entity& e : ents) std::printf(" ent=%.6f,%.6f,%.6f,%.6f,%d", e.x, e.y, e.vx, e.vy, e.ground);
        std::printf("\n");
    } catch (const std::exception& e) {
        std::fprintf(stderr, "host_demo: %s\n", e.what());
        return 1;
    }
    return 0;
}
