// drop_in_example.cpp -- the reference's doc-test (camera.rs:81-94) and bench (benches/simple_scene.rs:15-48)
// written against the C++ host mirror.  Build: `make examples` (needs a GPU to run).
#include <cstdio>
#include <memory>

#include "vanrijn.hpp"

using namespace vanrijn;

int main(int argc, char **argv) {
    // camera.rs:81-94: an empty scene renders tile by tile without error
    {
        Scene scene;
        scene.camera_location = Vec3(0.0, 0.0, 0.0);
        const size_t image_width = 640, image_height = 480;
        TileIterator tiles(image_width, image_height, 32);
        Tile tile;
        size_t n = 0;
        while (tiles.next(tile)) {
            AccumulationBuffer tile_image = partial_render_scene(scene, tile, image_height, image_width);
            n += tile_image.width() * tile_image.height();
        }
        std::printf("doc-test: %zu pixels rendered over an empty scene\n", n);
    }
    // benches/simple_scene.rs: 6x6, the bunny BVH with a reflective material
    if (argc > 1) {
        const size_t image_width = 6, image_height = 6;
        Scene scene;
        scene.camera_location = Vec3(-2.0, 1.0, -5.0);
        auto prims = load_obj(argv[1], std::make_shared<ReflectiveMaterial>(
                                           Spectrum::reflection_from_linear_rgb(ColourRgbF::from_named(NamedColour::Yellow)), 0.05, 0.9));
        scene.objects.push_back(BoundingVolumeHierarchy::build(prims));
        Tile tile{0, image_width, 0, image_height};
        AccumulationBuffer full(image_width, image_height);
        for (int pass = 0; pass < 16; pass++) { // main.rs:199-217: repeat 1-spp passes and merge
            RenderOptions o;
            o.sample_offset = pass;
            AccumulationBuffer b = partial_render_scene(scene, tile, image_height, image_width, o);
            full.merge_tile(tile, b);
        }
        std::printf("bench scene: pixel (3,3) XYZ = %g %g %g after %g samples\n", full.colour[3 * (3 * 6 + 3)],
                    full.colour[3 * (3 * 6 + 3) + 1], full.colour[3 * (3 * 6 + 3) + 2], full.weight[3 * 6 + 3]);
    }
    return 0;
}
