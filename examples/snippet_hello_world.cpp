// snippet_hello_world.cpp -- the reference's SnippetHelloWorld scene (physx/snippets/snippethelloworld/SnippetHelloWorld.cpp:64-79,
// createStack: stacks of boxes on a ground plane) stepped through the C++ host mirror (include/physx_b200.hpp) over the C ABI.
// BASELINE config 1: 10 stacks x 10 boxes (half-extent 0.5, zero gap), TGS 4+1 iterations, 60 Hz.
//   usage: snippet_hello_world [steps] [stacks] [height] [pgs]
// Prints one JSON line with a checksum of the final state (tests/test_gpu_parity.py compares it with the Python binding's result).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../include/physx_b200.hpp"

int main(int argc, char** argv) {
  const int steps = argc > 1 ? std::atoi(argv[1]) : 100, stacks = argc > 2 ? std::atoi(argv[2]) : 10, height = argc > 3 ? std::atoi(argv[3]) : 10;
  const bool pgs = argc > 4 && std::strcmp(argv[4], "pgs") == 0;
  try {
    const float he = 0.5f;
    std::vector<PxbActorRec> actors;
    actors.push_back(pxb::groundPlane());
    for (int s = 0; s < stacks; ++s)
      for (int j = 0; j < height; ++j) actors.push_back(pxb::dynamicBox(4.0f * s, he + 2.f * he * j, 0.f, he, he, he));
    pxb::Scene scene(pxb::defaultSceneDesc((uint32_t)actors.size(), pgs ? PXB_SOLVER_PGS : PXB_SOLVER_TGS));
    scene.addActors(actors);
    for (int i = 0; i < steps; ++i) { scene.simulate(1.0f / 60.0f); scene.fetchResults(true); }
    const std::vector<pxb::State> st = scene.getStates();
    double sum = 0, maxv = 0, miny = 1e9;
    for (const pxb::State& b : st) {
      for (int k = 0; k < 3; ++k) { sum += b.p[k]; if (std::fabs(b.linVel[k]) > maxv) maxv = std::fabs(b.linVel[k]); }
      for (int k = 0; k < 4; ++k) sum += b.q[k];
      if (b.p[1] < miny) miny = b.p[1];
    }
    std::printf("{\"bodies\": %u, \"steps\": %d, \"pairs\": %u, \"constraints\": %u, \"partitions\": %u, \"checksum\": %.9g, \"max_speed\": %.6g, \"min_y\": %.6g, \"top_y\": %.9g}\n",
                scene.getNbDynamics(), steps, scene.getNbPairs(), scene.getNbConstraints(), scene.getNbPartitions(), sum, maxv, miny, (double)st.back().p[1]);
    return 0;
  } catch (const pxb::Error& e) { std::fprintf(stderr, "%s\n", e.what()); return 1; }
}
