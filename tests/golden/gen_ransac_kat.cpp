// Regenerates the input cloud of the reference's RansacPlane.CalculateInlersPlane test
// (/root/reference/monolidar_fusion/test/test_monolidar_fusion.cpp:376-409): 18000 points on the
// plane n=(0,0,1), d=1.6, xy ~ U(-20,20), N(0,0.5) noise on x,y,z, std::default_random_engine
// seeded with 1234. Only <random> is needed, so the exact sequence of the reference's test is
// reproduced with the same libstdc++ distributions. Writes x y z intensity as float32 to stdout.
#include <cstdio>
#include <random>

int main() {
    int size_data = 18000;
    std::default_random_engine generator;
    generator.seed(1234);
    std::uniform_real_distribution<> distribution(-20., 20.);
    std::normal_distribution<float> noise_dist(0., 0.5);
    double nx = 0., ny = 0., nz = 1.;
    float d = 1.6;
    for (int i = 0; i < size_data; ++i) {
        float rand_num_x = distribution(generator);
        float rand_num_y = distribution(generator);
        float z = -(nx * rand_num_x + ny * rand_num_y + d) / nz;
        float p[4];
        p[0] = rand_num_x + noise_dist(generator);
        p[1] = rand_num_y + noise_dist(generator);
        p[2] = z + noise_dist(generator);
        p[3] = 150;
        fwrite(p, sizeof(float), 4, stdout);
    }
    return 0;
}
