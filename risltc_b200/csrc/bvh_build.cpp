// bvh_build.cpp -- host-side BVH construction (binned SAH, binary tree, children's
// boxes stored in the parent). Replaces the driver's BLAS/TLAS build that the
// reference requests in create_acceleration_structure (scene.c:142-406); like there,
// the input is the mesh dequantised with mul + add (scene.c:176-187).
#include "bvh_build.h"
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <future>
#include <limits>
#include <utility>

namespace {

struct Box {
	float lo[3], hi[3];
	void reset() { for (int k = 0; k != 3; ++k) { lo[k] = std::numeric_limits<float>::infinity(); hi[k] = -lo[k]; } }
	void grow(const Box& b) { for (int k = 0; k != 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
	void grow(const float* p) { for (int k = 0; k != 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
	float half_area() const {
		float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
		return (dx < 0.0f) ? 0.0f : dx * dy + dy * dz + dz * dx;
	}
};

struct Builder {
	const float* verts;            // 9 floats per triangle
	std::vector<Box> tri_box;
	std::vector<float> centroid;   // 3 per triangle
	std::vector<uint32_t> order;
	std::vector<BvhNodeHost>* nodes;
	float pad;
	static constexpr int kBins = 16;
	static constexpr int kParallelDepth = 4;           // 2^4 tasks at most
	static constexpr uint32_t kParallelCount = 100000; // subtrees below this many triangles are built by the task that reached them
	uint32_t kLeaf = 4;   // largest leaf

	// Returns the child reference for [first, first + count) and its bounds; inner nodes are appended to `out` in preorder
	// (a node, its left subtree, its right subtree). The two subtrees of the top levels of a large scene are built by two
	// tasks into vectors of their own and spliced in afterwards with their indices shifted: the ranges of `order` they work
	// on are disjoint, and the result is the very tree (and node order) of the sequential build -- 5 M triangles in ~2 s
	// instead of 7-8 s. The tasks end before build_bvh returns (no host thread outlives the call).
	int build(uint32_t first, uint32_t count, Box& bounds, std::vector<BvhNodeHost>& out, int depth = 0) {
		Box cb; cb.reset(); bounds.reset();
		for (uint32_t i = first; i != first + count; ++i) {
			bounds.grow(tri_box[order[i]]);
			cb.grow(&centroid[3 * (size_t) order[i]]);
		}
		if (count <= kLeaf) return ~(int) ((first << 4) | (count - 1));
		int axis = 0;
		float ext[3] = { cb.hi[0] - cb.lo[0], cb.hi[1] - cb.lo[1], cb.hi[2] - cb.lo[2] };
		if (ext[1] > ext[axis]) axis = 1;
		if (ext[2] > ext[axis]) axis = 2;
		uint32_t mid = first + count / 2;
		if (ext[axis] > 0.0f) {
			// binned SAH over the three axes
			float best_cost = std::numeric_limits<float>::infinity(); int best_axis = -1, best_bin = -1;
			for (int a = 0; a != 3; ++a) {
				if (!(ext[a] > 0.0f)) continue;
				Box bin_box[kBins]; uint32_t bin_n[kBins];
				for (int b = 0; b != kBins; ++b) { bin_box[b].reset(); bin_n[b] = 0; }
				float scale = (float) kBins / ext[a];
				for (uint32_t i = first; i != first + count; ++i) {
					int b = std::min(kBins - 1, (int) ((centroid[3 * (size_t) order[i] + a] - cb.lo[a]) * scale));
					bin_box[b].grow(tri_box[order[i]]); bin_n[b]++;
				}
				float right_area[kBins]; uint32_t right_n[kBins];
				Box acc; acc.reset(); uint32_t n = 0;
				for (int b = kBins - 1; b > 0; --b) { acc.grow(bin_box[b]); n += bin_n[b]; right_area[b] = acc.half_area(); right_n[b] = n; }
				acc.reset(); n = 0;
				for (int b = 0; b + 1 != kBins; ++b) {
					acc.grow(bin_box[b]); n += bin_n[b];
					if (n == 0 || right_n[b + 1] == 0) continue;
					float cost = acc.half_area() * (float) n + right_area[b + 1] * (float) right_n[b + 1];
					if (cost < best_cost) { best_cost = cost; best_axis = a; best_bin = b; }
				}
			}
			if (best_axis >= 0) {
				float scale = (float) kBins / ext[best_axis], lo = cb.lo[best_axis];
				auto it = std::partition(order.begin() + first, order.begin() + first + count, [&](uint32_t t) {
					int b = std::min(kBins - 1, (int) ((centroid[3 * (size_t) t + best_axis] - lo) * scale));
					return b <= best_bin;
				});
				mid = (uint32_t) (it - order.begin());
			}
			if (mid == first || mid == first + count) {
				mid = first + count / 2;
				std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count, [&](uint32_t x, uint32_t y) {
					return centroid[3 * (size_t) x + axis] < centroid[3 * (size_t) y + axis];
				});
			}
		}
		else if (count <= 16) return ~(int) ((first << 4) | (count - 1));  // coincident centroids: one fat leaf
		int index = (int) out.size();
		out.push_back(BvhNodeHost());
		Box lb, rb;
		int l, r;
		if (depth < kParallelDepth && count >= kParallelCount) {
			std::vector<BvhNodeHost> left_nodes, right_nodes;
			auto right_task = std::async(std::launch::async, [&]() { return build(mid, first + count - mid, rb, right_nodes, depth + 1); });
			l = build(first, mid - first, lb, left_nodes, depth + 1);
			r = right_task.get();
			const int left_base = (int) out.size(), right_base = left_base + (int) left_nodes.size();
			auto splice = [&](std::vector<BvhNodeHost>& part, int base) {
				for (BvhNodeHost& node : part) {
					if (node.left >= 0) node.left += base;
					if (node.right >= 0) node.right += base;
					out.push_back(node);
				}
			};
			splice(left_nodes, left_base);
			splice(right_nodes, right_base);
			if (l >= 0) l += left_base;
			if (r >= 0) r += right_base;
		}
		else {
			l = build(first, mid - first, lb, out, depth + 1);
			r = build(mid, first + count - mid, rb, out, depth + 1);
		}
		BvhNodeHost& n = out[index];
		for (int k = 0; k != 3; ++k) {
			n.left_lo[k] = lb.lo[k] - pad; n.left_hi[k] = lb.hi[k] + pad;
			n.right_lo[k] = rb.lo[k] - pad; n.right_hi[k] = rb.hi[k] + pad;
		}
		n.left = l; n.right = r;
		return index;
	}
};

}  // namespace

void dequantize_mesh_for_bvh(const uint32_t* q, uint64_t triangle_count, const float factor[3], const float summand[3], std::vector<float>& verts) {
	verts.resize(9 * triangle_count);
	for (uint64_t i = 0; i != triangle_count * 3; ++i) {
		uint32_t q0 = q[2 * i], q1 = q[2 * i + 1];
		float p[3] = {
			(float) (q0 & 0x1FFFFFu),
			(float) (((q0 & 0xFFE00000u) >> 21) | ((q1 & 0x3FFu) << 11)),
			(float) ((q1 & 0x7FFFFC00u) >> 10) };
		for (int j = 0; j != 3; ++j) {
			volatile float prod = p[j] * factor[j];   // mul then add, never contracted
			verts[3 * i + j] = prod + summand[j];
		}
	}
}

void build_bvh(const float* verts, uint64_t triangle_count, std::vector<BvhNodeHost>& nodes, std::vector<uint32_t>& order, uint32_t max_leaf) {
	Builder b;
	b.kLeaf = max_leaf < 1 ? 1 : (max_leaf > 16 ? 16 : max_leaf);
	b.verts = verts; b.nodes = &nodes;
	b.tri_box.resize(triangle_count); b.centroid.resize(3 * triangle_count); b.order.resize(triangle_count);
	Box all; all.reset();
	for (uint64_t t = 0; t != triangle_count; ++t) {
		Box bx; bx.reset();
		for (int k = 0; k != 3; ++k) bx.grow(verts + 9 * t + 3 * k);
		b.tri_box[t] = bx; all.grow(bx);
		for (int k = 0; k != 3; ++k) b.centroid[3 * t + k] = (verts[9 * t + k] + verts[9 * t + 3 + k] + verts[9 * t + 6 + k]) * (1.0f / 3.0f);
		b.order[t] = (uint32_t) t;
	}
	float extent = std::max(all.hi[0] - all.lo[0], std::max(all.hi[1] - all.lo[1], all.hi[2] - all.lo[2]));
	b.pad = 1.0e-5f * extent + 1.0e-7f;
	nodes.clear();
	nodes.reserve(triangle_count);
	Box root;
	// RISLTC_BVH_THREADS=1 builds on the calling thread alone (the tree is the same either way: tests/test_host_layer.py)
	const char* threads = getenv("RISLTC_BVH_THREADS");
	int ref = b.build(0, (uint32_t) triangle_count, root, nodes, (threads && atoi(threads) == 1) ? Builder::kParallelDepth : 0);
	if (ref < 0) {
		// the whole scene is one leaf: wrap it in a root whose right child can never be hit
		BvhNodeHost n;
		for (int k = 0; k != 3; ++k) {
			n.left_lo[k] = root.lo[k] - b.pad; n.left_hi[k] = root.hi[k] + b.pad;
			n.right_lo[k] = std::numeric_limits<float>::infinity(); n.right_hi[k] = -std::numeric_limits<float>::infinity();
		}
		n.left = ref; n.right = ref;
		nodes.push_back(n);
	}
	order.swap(b.order);
}

uint32_t bvh_depth(const std::vector<BvhNodeHost>& nodes) {
	if (nodes.empty()) return 0;
	uint32_t deepest = 0;
	std::vector<std::pair<int, uint32_t>> todo(1, std::make_pair(0, 1u));
	while (!todo.empty()) {
		const std::pair<int, uint32_t> it = todo.back(); todo.pop_back();
		deepest = std::max(deepest, it.second);
		const BvhNodeHost& n = nodes[it.first];
		if (n.left >= 0) todo.push_back(std::make_pair(n.left, it.second + 1u));
		if (n.right >= 0) todo.push_back(std::make_pair(n.right, it.second + 1u));
	}
	return deepest;
}

// ---- four children per node, 8-bit boxes
namespace {
struct Child4 { int ref; float lo[3], hi[3]; };

float half_area(const Child4& c) {
	float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
	return dx * dy + dy * dz + dz * dx;
}
void children_of(const BvhNodeHost& n, Child4& l, Child4& r) {
	l.ref = n.left; r.ref = n.right;
	for (int k = 0; k != 3; ++k) { l.lo[k] = n.left_lo[k]; l.hi[k] = n.left_hi[k]; r.lo[k] = n.right_lo[k]; r.hi[k] = n.right_hi[k]; }
}
}  // namespace

uint32_t build_qbvh4(const std::vector<BvhNodeHost>& binary, std::vector<Qbvh4NodeHost>& nodes4) {
	nodes4.clear();
	nodes4.reserve(binary.size() / 2 + 1);
	// work list of (binary node, slot of the 4-wide node that replaces it); children are emitted depth first so that a
	// subtree stays close in memory
	struct Item { int binary_node; uint32_t out, depth; };
	uint32_t max_depth = 1;
	std::vector<Item> todo;
	nodes4.push_back(Qbvh4NodeHost());
	todo.push_back({ 0, 0u, 1u });
	while (!todo.empty()) {
		Item it = todo.back(); todo.pop_back();
		max_depth = std::max(max_depth, it.depth);
		Child4 c[4]; int n = 2;
		children_of(binary[it.binary_node], c[0], c[1]);
		// the root of a one-leaf scene refers to the same leaf twice with an empty right box: keep only the left one
		if (!(c[1].lo[0] <= c[1].hi[0])) n = 1;
		while (n < 4) {
			int pick = -1; float best = -1.0f;
			for (int i = 0; i != n; ++i) if (c[i].ref >= 0 && half_area(c[i]) > best) { best = half_area(c[i]); pick = i; }
			if (pick < 0) break;
			Child4 l, r;
			children_of(binary[c[pick].ref], l, r);
			c[pick] = l; c[n++] = r;
		}
		double lo[3], hi[3];
		for (int k = 0; k != 3; ++k) {
			lo[k] = c[0].lo[k]; hi[k] = c[0].hi[k];
			for (int i = 1; i != n; ++i) { lo[k] = std::min(lo[k], (double) c[i].lo[k]); hi[k] = std::max(hi[k], (double) c[i].hi[k]); }
		}
		Qbvh4NodeHost q;
		memset(&q, 0, sizeof(q));
		const int scale_word[3] = { 3, 10, 11 };
		uint8_t qlo[3][4], qhi[3][4];
		for (int k = 0; k != 3; ++k) {
			double extent = std::max(hi[k] - lo[k], 1.0e-30);
			int e = (int) std::ceil(std::log2(extent / 250.0));
			for (;; ++e) {
				if (e < -100) e = -100;
				double s = std::ldexp(1.0, e);
				float origin = (float) (lo[k] - 2.0 * s);
				bool ok = true;
				for (int i = 0; i != n && ok; ++i) {
					// planes are evaluated on the device as origin + q * s with an error far below 0.05 steps
					double a = std::floor(((double) c[i].lo[k] - (double) origin) / s - 0.05);
					double b = std::ceil(((double) c[i].hi[k] - (double) origin) / s + 0.05);
					if (a < 0.0 || b > 255.0 || !((double) origin + a * s <= (double) c[i].lo[k]) || !((double) origin + b * s >= (double) c[i].hi[k])) ok = false;
					else { qlo[k][i] = (uint8_t) a; qhi[k][i] = (uint8_t) b; }
				}
				if (ok) {
					memcpy(&q.w[k], &origin, 4);
					float scale = (float) std::ldexp(1.0, e + 15);
					memcpy(&q.w[scale_word[k]], &scale, 4);
					break;
				}
			}
			for (int i = n; i != 4; ++i) { qlo[k][i] = 255; qhi[k][i] = 0; }   // inverted: never hit
		}
		for (int k = 0; k != 3; ++k)
			for (int i = 0; i != 4; ++i) {
				q.w[4 + k] |= (uint32_t) qlo[k][i] << (8 * i);
				q.w[7 + k] |= (uint32_t) qhi[k][i] << (8 * i);
			}
		for (int i = 0; i != 4; ++i) {
			int ref = 0x7FFFFFFF;
			if (i < n) {
				if (c[i].ref < 0) ref = c[i].ref;
				else {
					ref = (int) nodes4.size();
					nodes4.push_back(Qbvh4NodeHost());
					todo.push_back({ c[i].ref, (uint32_t) ref, it.depth + 1u });
				}
			}
			q.w[12 + i] = (uint32_t) ref;
		}
		nodes4[it.out] = q;
	}
	return max_depth;
}
