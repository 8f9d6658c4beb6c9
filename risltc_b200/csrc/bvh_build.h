// bvh_build.h -- host BVH builder interface (see bvh_build.cpp).
#pragma once
#include <stdint.h>
#include <vector>

struct BvhNodeHost {
	float left_lo[3], left_hi[3], right_lo[3], right_hi[3];
	int left, right;   // >= 0: inner node index; < 0: leaf, ~((first_slot << 4) | (count - 1))
};

// 21-bit positions -> world space exactly as scene.c:176-187 does for the acceleration structure.
void dequantize_mesh_for_bvh(const uint32_t* quantized_positions, uint64_t triangle_count,
	const float factor[3], const float summand[3], std::vector<float>& verts);

// verts: 9 floats per triangle. order[slot] = triangle index stored at that leaf slot.
void build_bvh(const float* verts, uint64_t triangle_count, std::vector<BvhNodeHost>& nodes, std::vector<uint32_t>& order, uint32_t max_leaf = 4);

// Levels of inner nodes of the binary tree (the per-thread traversal stacks hold one entry per level at most).
uint32_t bvh_depth(const std::vector<BvhNodeHost>& nodes);

// The binary tree collapsed to four children per node with 8-bit boxes (layout: Qbvh4Node in common.cuh, 16 words).
// Every quantised box contains the binary tree's (padded) box of the same child.
struct Qbvh4NodeHost { uint32_t w[16]; };
// Returns the depth of the 4-wide tree (levels of nodes).
uint32_t build_qbvh4(const std::vector<BvhNodeHost>& binary, std::vector<Qbvh4NodeHost>& nodes4);
