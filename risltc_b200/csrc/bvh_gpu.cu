// bvh_gpu.cu -- acceleration-structure build on the device (create_acceleration_structure, scene.c:142-406, asks the
// driver for a device-side build; this is its counterpart here, next to the host builder of bvh_build.cpp).
//
// Pipeline (all on the device, the mesh never returns to the host):
//   1. per triangle: dequantise the 21-bit positions exactly as scene.c:176-187 (mul, then add), 63-bit Morton key of the
//      quantised centroid (the quantisation grid already spans the mesh's bounding box: 21 bits per axis, no reduction);
//   2. radix sort of {key, triangle} (cub::DeviceRadixSort -- library code, a build step, not the frame path);
//   3. Karras' parallel binary radix tree over the sorted keys (one thread per inner node, duplicate keys split by index);
//   4. bottom-up boxes with one atomic ticket per inner node; subtrees of <= max_leaf triangles become leaves;
//   5. BvhNode / BvhTri records in the layout the traversal kernels read (children's boxes in the parent, padded);
//   6. the 4-wide, 8-bit collapse (Qbvh4Node), level by level: one thread per 4-wide node picks its children by largest
//      surface area and quantises their boxes outwards, exactly the rules of build_qbvh4 (bvh_build.cpp).
// Hit / no-hit decisions do not depend on the shape of the tree (bvh.cuh), so images are bit-identical to those rendered
// with the host-built tree (tests/test_gpu_frames.py); what differs is traversal cost: a Morton-order tree is ~1.2-1.6x
// more expensive to traverse than the binned-SAH tree of the host builder, and ~100x faster to build.
#define RL_NS gpubuild
#include "internal.h"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <cub/device/device_scan.cuh>
#include <vector>

using namespace RL_NS;

namespace {

struct Box6 { float lo[3], hi[3]; };

__device__ __forceinline__ void dequantize(uint2 q, const float* factor, const float* summand, float* out) {
	const float p[3] = {
		(float) (q.x & 0x1FFFFFu),
		(float) (((q.x & 0xFFE00000u) >> 21) | ((q.y & 0x3FFu) << 11)),
		(float) ((q.y & 0x7FFFFC00u) >> 10) };
	#pragma unroll
	for (int j = 0; j != 3; ++j) out[j] = __fadd_rn(__fmul_rn(p[j], factor[j]), summand[j]);
}
__device__ __forceinline__ void quantized_xyz(uint2 q, uint32_t* out) {
	out[0] = q.x & 0x1FFFFFu;
	out[1] = ((q.x & 0xFFE00000u) >> 21) | ((q.y & 0x3FFu) << 11);
	out[2] = (q.y & 0x7FFFFC00u) >> 10;
}
// 21 bits -> every third bit of 63
__device__ __forceinline__ unsigned long long spread21(uint32_t v) {
	unsigned long long x = v & 0x1FFFFFull;
	x = (x | x << 32) & 0x1F00000000FFFFull;
	x = (x | x << 16) & 0x1F0000FF0000FFull;
	x = (x | x << 8) & 0x100F00F00F00F00Full;
	x = (x | x << 4) & 0x10C30C30C30C30C3ull;
	x = (x | x << 2) & 0x1249249249249249ull;
	return x;
}

struct DequantParams { float factor[3], summand[3]; };

__global__ void keys_kernel(const uint2* positions, uint32_t T, unsigned long long* keys, uint32_t* values) {
	const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= T) return;
	uint32_t a[3], b[3], c[3];
	quantized_xyz(positions[3 * (size_t) t], a); quantized_xyz(positions[3 * (size_t) t + 1], b); quantized_xyz(positions[3 * (size_t) t + 2], c);
	keys[t] = spread21((a[0] + b[0] + c[0]) / 3u) << 2 | spread21((a[1] + b[1] + c[1]) / 3u) << 1 | spread21((a[2] + b[2] + c[2]) / 3u);
	values[t] = t;
}

// Triangle records in sorted order + their boxes
__global__ void tris_kernel(const uint2* positions, const uint32_t* order, uint32_t T, DequantParams dq, BvhTri* tris, Box6* leaf_box) {
	const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= T) return;
	const uint32_t t = order[slot];
	const uint2 q0 = positions[3 * (size_t) t], q1 = positions[3 * (size_t) t + 1], q2 = positions[3 * (size_t) t + 2];
	float v0[3], v1[3], v2[3];
	dequantize(q0, dq.factor, dq.summand, v0); dequantize(q1, dq.factor, dq.summand, v1); dequantize(q2, dq.factor, dq.summand, v2);
	const uint32_t id = t | ((q0.y >> 31) << 31);
	BvhTri r;
	r.v0 = make_float4(v0[0], v0[1], v0[2], __uint_as_float(id));
	r.e1 = make_float4(__fsub_rn(v1[0], v0[0]), __fsub_rn(v1[1], v0[1]), __fsub_rn(v1[2], v0[2]), 0.0f);
	r.e2 = make_float4(__fsub_rn(v2[0], v0[0]), __fsub_rn(v2[1], v0[1]), __fsub_rn(v2[2], v0[2]), 0.0f);
	if (tris) tris[slot] = r;
	Box6 b;
	#pragma unroll
	for (int k = 0; k != 3; ++k) { b.lo[k] = fminf(v0[k], fminf(v1[k], v2[k])); b.hi[k] = fmaxf(v0[k], fmaxf(v1[k], v2[k])); }
	if (leaf_box) leaf_box[slot] = b;
}

// length of the common prefix of keys i and j (duplicates: continue with the index), -1 outside the array
__device__ __forceinline__ int prefix(const unsigned long long* keys, int T, int i, int j) {
	if (j < 0 || j >= T) return -1;
	const unsigned long long x = keys[i] ^ keys[j];
	return x ? __clzll((long long) x) : 64 + __clz(i ^ j);
}

// Karras 2012: inner node i covers [first, last]; children are leaves (index | 0x80000000) or inner nodes
__global__ void hierarchy_kernel(const unsigned long long* keys, int T, int2* children, int2* range, int* inner_parent, int* leaf_parent) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= T - 1) return;
	const int d = (prefix(keys, T, i, i + 1) - prefix(keys, T, i, i - 1)) >= 0 ? 1 : -1;
	const int min_prefix = prefix(keys, T, i, i - d);
	int l_max = 2;
	while (prefix(keys, T, i, i + l_max * d) > min_prefix) l_max *= 2;
	int l = 0;
	for (int t = l_max / 2; t >= 1; t /= 2) if (prefix(keys, T, i, i + (l + t) * d) > min_prefix) l += t;
	const int j = i + l * d;
	const int node_prefix = prefix(keys, T, i, j);
	int s = 0;
	for (int t = (l + 1) / 2;; t = (t + 1) / 2) {
		if (prefix(keys, T, i, i + (s + t) * d) > node_prefix) s += t;
		if (t == 1) break;
	}
	const int gamma = i + s * d + min(d, 0);
	const int first = min(i, j), last = max(i, j);
	const bool left_leaf = first == gamma, right_leaf = last == gamma + 1;
	children[i] = make_int2(left_leaf ? (gamma | (int) 0x80000000) : gamma, right_leaf ? ((gamma + 1) | (int) 0x80000000) : gamma + 1);
	range[i] = make_int2(first, last);
	if (left_leaf) leaf_parent[gamma] = i; else inner_parent[gamma] = i;
	if (right_leaf) leaf_parent[gamma + 1] = i; else inner_parent[gamma + 1] = i;
	if (i == 0) inner_parent[0] = -1;
}

__device__ __forceinline__ Box6 merge(const Box6& a, const Box6& b) {
	Box6 r;
	#pragma unroll
	for (int k = 0; k != 3; ++k) { r.lo[k] = fminf(a.lo[k], b.lo[k]); r.hi[k] = fmaxf(a.hi[k], b.hi[k]); }
	return r;
}

// One thread per leaf walks towards the root; the second thread to arrive at a node computes its box and goes on
// (boxes written by other SMs are read past the non-coherent L1)
__device__ __forceinline__ Box6 load_box_cg(const Box6* p) {
	Box6 b;
	const float* f = (const float*) p;
	#pragma unroll
	for (int k = 0; k != 3; ++k) { b.lo[k] = __ldcg(f + k); b.hi[k] = __ldcg(f + 3 + k); }
	return b;
}
__global__ void boxes_kernel(int T, const int2* children, const int* inner_parent, const int* leaf_parent, const Box6* leaf_box, Box6* node_box, unsigned int* arrived) {
	const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
	if (leaf >= T) return;
	int node = leaf_parent[leaf];
	while (node >= 0) {
		__threadfence();
		if (atomicAdd(&arrived[node], 1u) == 0u) return;
		__threadfence();
		const int2 c = children[node];
		const Box6 l = (c.x < 0) ? leaf_box[c.x & 0x7FFFFFFF] : load_box_cg(node_box + c.x);
		const Box6 r = (c.y < 0) ? leaf_box[c.y & 0x7FFFFFFF] : load_box_cg(node_box + c.y);
		node_box[node] = merge(l, r);
		node = inner_parent[node];
	}
}

// Levels of the tree: every leaf counts its ancestors (the walk above only measures one path per node)
__global__ void depth_kernel(int T, const int* inner_parent, const int* leaf_parent, unsigned int* depth) {
	const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
	if (leaf >= T) return;
	unsigned int levels = 0;
	for (int node = leaf_parent[leaf]; node >= 0; node = inner_parent[node]) ++levels;
	levels = __reduce_max_sync(__activemask(), levels);
	if ((threadIdx.x & 31u) == 0u) atomicMax(depth, levels);
}

// BvhNode records: a child that covers <= max_leaf triangles is a leaf reference ~((first << 4) | (count - 1))
__global__ void nodes_kernel(int T, uint32_t max_leaf, float pad, const int2* children, const int2* range, const Box6* leaf_box, const Box6* node_box, BvhNode* nodes) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= T - 1) return;
	const int2 own = range[i];
	if ((uint32_t) (own.y - own.x + 1) <= max_leaf && i != 0) return;   // inside a leaf: never referenced
	const int2 c = children[i];
	int ref[2]; Box6 box[2];
	#pragma unroll
	for (int side = 0; side != 2; ++side) {
		const int child = side ? c.y : c.x;
		if (child < 0) {
			const int slot = child & 0x7FFFFFFF;
			ref[side] = ~(int) (((uint32_t) slot << 4) | 0u);
			box[side] = leaf_box[slot];
		}
		else {
			const int2 r = range[child];
			const uint32_t count = (uint32_t) (r.y - r.x + 1);
			ref[side] = (count <= max_leaf) ? ~(int) (((uint32_t) r.x << 4) | (count - 1u)) : child;
			box[side] = node_box[child];
		}
	}
	BvhNode n;
	n.a = make_float4(box[0].lo[0] - pad, box[0].lo[1] - pad, box[0].lo[2] - pad, box[0].hi[0] + pad);
	n.b = make_float4(box[0].hi[1] + pad, box[0].hi[2] + pad, box[1].lo[0] - pad, box[1].lo[1] - pad);
	n.c = make_float4(box[1].lo[2] - pad, box[1].hi[0] + pad, box[1].hi[1] + pad, box[1].hi[2] + pad);
	n.d = make_int4(ref[0], ref[1], 0, 0);
	nodes[i] = n;
}

// ---- PLOC (parallel locally-ordered clustering, Meister & Bittner 2018): bottom-up agglomeration along the Morton curve.
// Clusters ids: 0 .. T-1 are the triangles in Morton order, T .. 2T-2 the inner nodes in the order they are created (the
// root last). Every round each cluster looks at its `radius` neighbours on either side for the partner with the smallest
// merged surface area; mutual choices merge. Trees come out within a few per cent of the binned-SAH builder's traversal cost.
#define RL_PLOC_BLOCK 256
#define RL_PLOC_MAX_RADIUS 32
#define RL_PLOC_NONE 0xFFFFFFFFu
__device__ __forceinline__ float box_half_area(const Box6& a, const Box6& b) {
	const float dx = fmaxf(a.hi[0], b.hi[0]) - fminf(a.lo[0], b.lo[0]), dy = fmaxf(a.hi[1], b.hi[1]) - fminf(a.lo[1], b.lo[1]), dz = fmaxf(a.hi[2], b.hi[2]) - fminf(a.lo[2], b.lo[2]);
	return dx * dy + dy * dz + dz * dx;
}
__global__ void __launch_bounds__(RL_PLOC_BLOCK) ploc_neighbour_kernel(const uint32_t* cluster, int n, const Box6* box, int* nearest, int radius) {
	__shared__ Box6 tile[RL_PLOC_BLOCK + 2 * RL_PLOC_MAX_RADIUS];
	const int base = (int) (blockIdx.x * RL_PLOC_BLOCK) - radius;
	for (int k = threadIdx.x; k < RL_PLOC_BLOCK + 2 * radius; k += RL_PLOC_BLOCK) {
		const int j = base + k;
		if (j >= 0 && j < n) tile[k] = box[cluster[j]];
	}
	__syncthreads();
	const int i = (int) (blockIdx.x * RL_PLOC_BLOCK + threadIdx.x);
	if (i >= n) return;
	const Box6 mine = tile[threadIdx.x + radius];
	float best = INFINITY; int best_j = -1, best_distance = 0x7FFFFFFF;
	// candidates in the order: distance 1, 2, ...; at equal distance the side that pairs (0,1), (2,3), ... first, so that equal
	// costs (coincident geometry) still produce mutual choices and halve the cluster count every round
	for (int distance = 1; distance <= radius; ++distance) {
		#pragma unroll
		for (int side = 0; side != 2; ++side) {
			const bool forward = ((i & 1) == 0) == (side == 0);
			const int j = forward ? i + distance : i - distance;
			if (j < 0 || j >= n) continue;
			const float area = box_half_area(mine, tile[j - base]);
			if (area < best || (area == best && distance < best_distance)) { best = area; best_j = j; best_distance = distance; }
		}
	}
	nearest[i] = best_j;
}
__global__ void ploc_merge_kernel(const uint32_t* cluster, int n, const int* nearest, Box6* box, int2* kids, int* parent, unsigned int* node_count, uint32_t* out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int j = nearest[i];
	const uint32_t mine = cluster[i];
	if (j < 0 || nearest[j] != i) { out[i] = mine; return; }
	if (i > j) { out[i] = RL_PLOC_NONE; return; }
	const uint32_t other = cluster[j];
	const uint32_t id = atomicAdd(node_count, 1u);
	kids[id] = make_int2((int) mine, (int) other);
	parent[mine] = (int) id; parent[other] = (int) id;
	box[id] = merge(box[mine], box[other]);
	out[i] = id;
}
struct IsCluster { __device__ __forceinline__ bool operator()(const uint32_t& c) const { return c != RL_PLOC_NONE; } };

// Leaves of two triangles: an inner node whose children are both triangles becomes one leaf, and its triangles get adjacent
// slots. unit[s] = slots that triangle s (Morton order) claims at its place of the new order: 2 for the first child of such a
// node (itself + its sibling), 0 for the second, 1 for a triangle that stays a leaf of its own.
__global__ void ploc_units_kernel(int T, uint32_t max_leaf, const int* parent, const int2* kids, uint32_t* unit) {
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= T) return;
	const int2 k = kids[parent[s]];
	const bool pair = max_leaf >= 2u && k.x < T && k.y < T;
	unit[s] = pair ? (s == k.x ? 2u : 0u) : 1u;
}
__global__ void ploc_slots_kernel(int T, const uint32_t* unit, const uint32_t* base, const int* parent, const int2* kids, const uint32_t* order, uint32_t* new_slot, uint32_t* new_order) {
	const int s = blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= T) return;
	const uint32_t slot = (unit[s] == 0u) ? base[kids[parent[s]].x] + 1u : base[s];
	new_slot[s] = slot;
	new_order[slot] = order[s];
}
// BvhNode records; inner id -> index (last - id), so that the root (created last) is node 0
__global__ void ploc_nodes_kernel(int T, int last, uint32_t max_leaf, float pad, const int2* kids, const Box6* box, const uint32_t* new_slot, BvhNode* nodes) {
	const int id = T + blockIdx.x * blockDim.x + threadIdx.x;
	if (id > last) return;
	const int2 k = kids[id];
	if (max_leaf >= 2u && k.x < T && k.y < T && id != last) return;   // a two-triangle leaf: referenced as a leaf by its parent
	int ref[2];
	#pragma unroll
	for (int side = 0; side != 2; ++side) {
		const int c = side ? k.y : k.x;
		if (c < T) ref[side] = ~(int) ((new_slot[c] << 4) | 0u);
		else {
			const int2 g = kids[c];
			ref[side] = (max_leaf >= 2u && g.x < T && g.y < T) ? ~(int) ((new_slot[g.x] << 4) | 1u) : last - c;
		}
	}
	const Box6 l = box[k.x], r = box[k.y];
	BvhNode n;
	n.a = make_float4(l.lo[0] - pad, l.lo[1] - pad, l.lo[2] - pad, l.hi[0] + pad);
	n.b = make_float4(l.hi[1] + pad, l.hi[2] + pad, r.lo[0] - pad, r.lo[1] - pad);
	n.c = make_float4(r.lo[2] - pad, r.hi[0] + pad, r.hi[1] + pad, r.hi[2] + pad);
	n.d = make_int4(ref[0], ref[1], 0, 0);
	nodes[last - id] = n;
}
__global__ void ploc_depth_kernel(int T, const int* parent, unsigned int* depth) {
	const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
	if (leaf >= T) return;
	unsigned int levels = 0;
	for (int node = parent[leaf]; node >= 0; node = parent[node]) ++levels;
	levels = __reduce_max_sync(__activemask(), levels);
	if ((threadIdx.x & 31u) == 0u) atomicMax(depth, levels);
}

// ---- 4-wide collapse, one level per launch
struct Child4 { int ref; float lo[3], hi[3]; };
__device__ __forceinline__ float half_area(const Child4& c) {
	const float dx = c.hi[0] - c.lo[0], dy = c.hi[1] - c.lo[1], dz = c.hi[2] - c.lo[2];
	return dx * dy + dy * dz + dz * dx;
}
__device__ __forceinline__ void children_of(const BvhNode& n, Child4& l, Child4& r) {
	l.ref = n.d.x; r.ref = n.d.y;
	l.lo[0] = n.a.x; l.lo[1] = n.a.y; l.lo[2] = n.a.z; l.hi[0] = n.a.w; l.hi[1] = n.b.x; l.hi[2] = n.b.y;
	r.lo[0] = n.b.z; r.lo[1] = n.b.w; r.lo[2] = n.c.x; r.hi[0] = n.c.y; r.hi[1] = n.c.z; r.hi[2] = n.c.w;
}
struct WideItem { int binary_node; uint32_t out; };

__global__ void collapse_kernel(const BvhNode* binary, const WideItem* in, uint32_t in_count, WideItem* out, unsigned int* out_count, unsigned int* node_count, Qbvh4Node* nodes4) {
	const uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
	if (item >= in_count) return;
	const WideItem it = in[item];
	Child4 c[4]; int n = 2;
	children_of(binary[it.binary_node], c[0], c[1]);
	while (n < 4) {
		int pick = -1; float best = -1.0f;
		for (int i = 0; i != n; ++i) if (c[i].ref >= 0 && half_area(c[i]) > best) { best = half_area(c[i]); pick = i; }
		if (pick < 0) break;
		Child4 l, r;
		children_of(binary[c[pick].ref], l, r);
		c[pick] = l; c[n++] = r;
	}
	uint32_t w[16];
	#pragma unroll
	for (int i = 0; i != 16; ++i) w[i] = 0u;
	const int scale_word[3] = { 3, 10, 11 };
	for (int k = 0; k != 3; ++k) {
		double lo = c[0].lo[k], hi = c[0].hi[k];
		for (int i = 1; i != n; ++i) { lo = fmin(lo, (double) c[i].lo[k]); hi = fmax(hi, (double) c[i].hi[k]); }
		const double extent = fmax(hi - lo, 1.0e-30);
		int e = (int) ceil(log2(extent / 250.0));
		uint32_t qlo[4], qhi[4];
		for (;; ++e) {
			if (e < -100) e = -100;
			const double s = ldexp(1.0, e);
			const float origin = (float) (lo - 2.0 * s);
			bool ok = true;
			for (int i = 0; i != n && ok; ++i) {
				// the same outward rounding and range test as build_qbvh4 (bvh_build.cpp)
				const double a = floor(((double) c[i].lo[k] - (double) origin) / s - 0.05);
				const double b = ceil(((double) c[i].hi[k] - (double) origin) / s + 0.05);
				if (a < 0.0 || b > 255.0 || !((double) origin + a * s <= (double) c[i].lo[k]) || !((double) origin + b * s >= (double) c[i].hi[k])) ok = false;
				else { qlo[i] = (uint32_t) a; qhi[i] = (uint32_t) b; }
			}
			if (ok) {
				w[k] = __float_as_uint(origin);
				w[scale_word[k]] = __float_as_uint((float) ldexp(1.0, e + 15));
				break;
			}
		}
		for (int i = n; i != 4; ++i) { qlo[i] = 255u; qhi[i] = 0u; }   // inverted: never hit
		for (int i = 0; i != 4; ++i) { w[4 + k] |= qlo[i] << (8 * i); w[7 + k] |= qhi[i] << (8 * i); }
	}
	uint32_t inner = 0;
	for (int i = 0; i != n; ++i) inner += c[i].ref >= 0;
	uint32_t slot = inner ? atomicAdd(node_count, inner) : 0u, queue = inner ? atomicAdd(out_count, inner) : 0u;
	for (int i = 0; i != 4; ++i) {
		int ref = RL_Q4_EMPTY;
		if (i < n) {
			if (c[i].ref < 0) ref = c[i].ref;
			else {
				ref = (int) slot;
				out[queue].binary_node = c[i].ref; out[queue].out = slot;
				++slot; ++queue;
			}
		}
		w[12 + i] = (uint32_t) ref;
	}
	Qbvh4Node q;
	q.a = make_uint4(w[0], w[1], w[2], w[3]); q.b = make_uint4(w[4], w[5], w[6], w[7]); q.c = make_uint4(w[8], w[9], w[10], w[11]);
	q.refs = make_int4((int) w[12], (int) w[13], (int) w[14], (int) w[15]);
	nodes4[it.out] = q;
}

template <class T> struct Scratch {
	T* p = nullptr;
	cudaError_t alloc(size_t n) { return cudaMalloc(&p, sizeof(T) * (n ? n : 1)); }
	T* release() { T* r = p; p = nullptr; return r; }
	~Scratch() { cudaFree(p); }
};
struct Events {
	cudaEvent_t e[4] = { nullptr, nullptr, nullptr, nullptr };
	cudaError_t create() { for (auto& x : e) { cudaError_t r = cudaEventCreate(&x); if (r != cudaSuccess) return r; } return cudaSuccess; }
	~Events() { for (auto& x : e) if (x) cudaEventDestroy(x); }
};

}  // namespace

// Builds the three arrays of SceneView from the quantised positions already on the device. On success the caller owns
// *nodes, *tris, *nodes4 (cudaFree). ms[0..2] = sort + hierarchy, boxes + records, 4-wide collapse (device time).
__global__ void iota_kernel(uint32_t* v, int n) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) v[i] = (uint32_t) i; }

// ploc_radius: 0 = Karras' radix tree over the Morton order (fastest build), 1..32 = PLOC with that search radius
int rl_build_bvh_gpu(const uint2* positions, uint64_t triangle_count, const float factor[3], const float summand[3], uint32_t max_leaf, uint32_t ploc_radius,
	BvhNode** nodes, BvhTri** tris, Qbvh4Node** nodes4, uint64_t counts[2], uint32_t depths[2], float ms[3])
{
	const int T = (int) triangle_count;
	if (max_leaf < 1) max_leaf = 1;
	if (max_leaf > 16) max_leaf = 16;
	if (triangle_count <= max_leaf || triangle_count < 2) return rl_fail("build_bvh_gpu: too few triangles for the device builder", nullptr);
	DequantParams dq;
	float extent = 0.0f;
	for (int k = 0; k != 3; ++k) { dq.factor[k] = factor[k]; dq.summand[k] = summand[k]; extent = fmaxf(extent, fabsf(factor[k]) * 2097151.0f); }
	const float pad = 1.0e-5f * extent + 1.0e-7f;
	Events events;   // everything allocated here is released on every return path (Scratch / Events destructors)
	CU(events.create());
	cudaEvent_t* ev = events.e;
	Scratch<unsigned long long> keys, keys_sorted; Scratch<uint32_t> values, order; Scratch<unsigned char> temp;
	Scratch<int2> children, range; Scratch<int> inner_parent, leaf_parent; Scratch<Box6> leaf_box, node_box; Scratch<unsigned int> arrived, scalars;
	Scratch<WideItem> frontier[2];
	CU(keys.alloc(T)); CU(keys_sorted.alloc(T)); CU(values.alloc(T)); CU(order.alloc(T));
	CU(children.alloc(T)); CU(range.alloc(T)); CU(inner_parent.alloc(T)); CU(leaf_parent.alloc(T)); CU(leaf_box.alloc(T)); CU(node_box.alloc(T));
	CU(arrived.alloc(T)); CU(scalars.alloc(4));
	CU(cudaMemset(arrived.p, 0, sizeof(unsigned int) * T)); CU(cudaMemset(scalars.p, 0, 4 * sizeof(unsigned int)));
	Scratch<BvhNode> nodes_out; Scratch<BvhTri> tris_out; Scratch<Qbvh4Node> wide_all, wide_compact;
	CU(nodes_out.alloc((size_t) (T - 1)));
	CU(tris_out.alloc((size_t) T));
	BvhNode* const out_nodes = nodes_out.p; BvhTri* const out_tris = tris_out.p;
	CU(cudaMemset(out_nodes, 0, sizeof(BvhNode) * (size_t) (T - 1)));   // slots of nodes inside leaves stay unreferenced
	const int block = 256, grid = (T + block - 1) / block;
	CU(cudaEventRecord(ev[0]));
	keys_kernel<<<grid, block>>>(positions, (uint32_t) T, keys.p, values.p);
	size_t temp_bytes = 0;
	CU(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys.p, keys_sorted.p, values.p, order.p, T, 0, 63));
	CU(temp.alloc(temp_bytes));
	CU(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, keys.p, keys_sorted.p, values.p, order.p, T, 0, 63));
	if (ploc_radius > RL_PLOC_MAX_RADIUS) ploc_radius = RL_PLOC_MAX_RADIUS;
	if (T < 4) ploc_radius = 0;
	if (ploc_radius) {
		Scratch<Box6> box; Scratch<int2> kids; Scratch<int> parent, nearest; Scratch<uint32_t> cluster[2], unit, base, new_slot, new_order; Scratch<unsigned char> temp2;
		CU(box.alloc(2 * (size_t) T)); CU(kids.alloc(2 * (size_t) T)); CU(parent.alloc(2 * (size_t) T)); CU(nearest.alloc(T));
		CU(cluster[0].alloc(T)); CU(cluster[1].alloc(T)); CU(unit.alloc(T)); CU(base.alloc(T)); CU(new_slot.alloc(T)); CU(new_order.alloc(T));
		size_t select_bytes = 0, scan_bytes = 0;
		CU(cub::DeviceSelect::If(nullptr, select_bytes, cluster[0].p, cluster[1].p, scalars.p + 1, T, IsCluster()));
		CU(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, unit.p, base.p, T));
		CU(temp2.alloc(select_bytes > scan_bytes ? select_bytes : scan_bytes));
		size_t temp2_bytes = select_bytes > scan_bytes ? select_bytes : scan_bytes;
		tris_kernel<<<grid, block>>>(positions, order.p, (uint32_t) T, dq, nullptr, box.p);      // boxes of the triangles in Morton order
		iota_kernel<<<grid, block>>>(cluster[0].p, T);
		const unsigned int first_inner = (unsigned int) T;
		CU(cudaMemcpy(scalars.p + 3, &first_inner, sizeof(first_inner), cudaMemcpyHostToDevice));   // scalars[3]: next node id
		int n = T, rounds = 0, side = 0;
		while (n > 1) {
			const int g = (n + RL_PLOC_BLOCK - 1) / RL_PLOC_BLOCK;
			ploc_neighbour_kernel<<<g, RL_PLOC_BLOCK>>>(cluster[side].p, n, box.p, nearest.p, (int) ploc_radius);
			ploc_merge_kernel<<<g, RL_PLOC_BLOCK>>>(cluster[side].p, n, nearest.p, box.p, kids.p, parent.p, scalars.p + 3, cluster[side ^ 1].p);
			// compact the survivors and the new nodes in place of their left members (the order along the curve is kept)
			CU(cub::DeviceSelect::If(temp2.p, temp2_bytes, cluster[side ^ 1].p, cluster[side].p, scalars.p + 1, n, IsCluster()));
			unsigned int survivors = 0;
			CU(cudaMemcpy(&survivors, scalars.p + 1, sizeof(survivors), cudaMemcpyDeviceToHost));
			if ((int) survivors >= n || survivors == 0u) return rl_fail("build_bvh_gpu: clustering made no progress", nullptr);
			n = (int) survivors;
			if (++rounds > 4096) return rl_fail("build_bvh_gpu: clustering does not terminate", nullptr);
		}
		const int last = 2 * T - 2;   // the root: T - 1 inner nodes were created, the last one by the final merge
		const int no_parent = -1;
		CU(cudaMemcpy(parent.p + last, &no_parent, sizeof(int), cudaMemcpyHostToDevice));
		CU(cudaEventRecord(ev[1]));
		ploc_units_kernel<<<grid, block>>>(T, max_leaf, parent.p, kids.p, unit.p);
		CU(cub::DeviceScan::ExclusiveSum(temp2.p, temp2_bytes, unit.p, base.p, T));
		ploc_slots_kernel<<<grid, block>>>(T, unit.p, base.p, parent.p, kids.p, order.p, new_slot.p, new_order.p);
		tris_kernel<<<grid, block>>>(positions, new_order.p, (uint32_t) T, dq, out_tris, nullptr);
		ploc_nodes_kernel<<<(T - 1 + block - 1) / block, block>>>(T, last, max_leaf, pad, kids.p, box.p, new_slot.p, out_nodes);
		ploc_depth_kernel<<<grid, block>>>(T, parent.p, scalars.p);
		CU(cudaEventRecord(ev[2]));
		CU(cudaDeviceSynchronize());   // the scratch arrays of this block are freed when it ends
	}
	else {
		hierarchy_kernel<<<grid, block>>>(keys_sorted.p, T, children.p, range.p, inner_parent.p, leaf_parent.p);
		CU(cudaEventRecord(ev[1]));
		tris_kernel<<<grid, block>>>(positions, order.p, (uint32_t) T, dq, out_tris, leaf_box.p);
		boxes_kernel<<<grid, block>>>(T, children.p, inner_parent.p, leaf_parent.p, leaf_box.p, node_box.p, arrived.p);
		depth_kernel<<<grid, block>>>(T, inner_parent.p, leaf_parent.p, scalars.p);
		nodes_kernel<<<grid, block>>>(T, max_leaf, pad, children.p, range.p, leaf_box.p, node_box.p, out_nodes);
		CU(cudaEventRecord(ev[2]));
	}
	// 4-wide collapse: frontier of (binary node, output slot), one launch per level
	CU(wide_all.alloc((size_t) T));
	Qbvh4Node* const wide = wide_all.p;
	CU(frontier[0].alloc(T)); CU(frontier[1].alloc(T));
	const WideItem root = { 0, 0u };
	CU(cudaMemcpy(frontier[0].p, &root, sizeof(root), cudaMemcpyHostToDevice));
	unsigned int one = 1u;
	CU(cudaMemcpy(scalars.p + 2, &one, sizeof(one), cudaMemcpyHostToDevice));   // scalars[2]: 4-wide nodes so far
	uint32_t level_count = 1, levels = 0;
	while (level_count) {
		CU(cudaMemset(scalars.p + 1, 0, sizeof(unsigned int)));                  // scalars[1]: items of the next level
		collapse_kernel<<<(level_count + 127) / 128, 128>>>(out_nodes, frontier[levels & 1].p, level_count, frontier[(levels + 1) & 1].p, scalars.p + 1, scalars.p + 2, wide);
		CU(cudaMemcpy(&level_count, scalars.p + 1, sizeof(unsigned int), cudaMemcpyDeviceToHost));
		++levels;
		if (levels > 200) return rl_fail("build_bvh_gpu: the 4-wide collapse does not terminate", nullptr);
	}
	CU(cudaEventRecord(ev[3]));
	CU(cudaDeviceSynchronize());
	CU(cudaGetLastError());
	unsigned int host_scalars[4];
	CU(cudaMemcpy(host_scalars, scalars.p, sizeof(host_scalars), cudaMemcpyDeviceToHost));
	// shrink the 4-wide array to its size
	CU(wide_compact.alloc((size_t) host_scalars[2]));
	CU(cudaMemcpy(wide_compact.p, wide, sizeof(Qbvh4Node) * (size_t) host_scalars[2], cudaMemcpyDeviceToDevice));
	for (int i = 0; i != 3; ++i) CU(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
	*nodes = nodes_out.release(); *tris = tris_out.release(); *nodes4 = wide_compact.release();
	counts[0] = (uint64_t) (T - 1); counts[1] = host_scalars[2];
	depths[0] = host_scalars[0]; depths[1] = levels;
	return 0;
}
