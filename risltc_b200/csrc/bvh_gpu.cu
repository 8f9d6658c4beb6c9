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
	tris[slot] = r;
	Box6 b;
	#pragma unroll
	for (int k = 0; k != 3; ++k) { b.lo[k] = fminf(v0[k], fminf(v1[k], v2[k])); b.hi[k] = fmaxf(v0[k], fmaxf(v1[k], v2[k])); }
	leaf_box[slot] = b;
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
	~Scratch() { cudaFree(p); }
};

}  // namespace

// Builds the three arrays of SceneView from the quantised positions already on the device. On success the caller owns
// *nodes, *tris, *nodes4 (cudaFree). ms[0..2] = sort + hierarchy, boxes + records, 4-wide collapse (device time).
int rl_build_bvh_gpu(const uint2* positions, uint64_t triangle_count, const float factor[3], const float summand[3], uint32_t max_leaf,
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
	cudaEvent_t ev[4];
	for (auto& e : ev) CU(cudaEventCreate(&e));
	Scratch<unsigned long long> keys, keys_sorted; Scratch<uint32_t> values, order; Scratch<unsigned char> temp;
	Scratch<int2> children, range; Scratch<int> inner_parent, leaf_parent; Scratch<Box6> leaf_box, node_box; Scratch<unsigned int> arrived, scalars;
	Scratch<WideItem> frontier[2];
	CU(keys.alloc(T)); CU(keys_sorted.alloc(T)); CU(values.alloc(T)); CU(order.alloc(T));
	CU(children.alloc(T)); CU(range.alloc(T)); CU(inner_parent.alloc(T)); CU(leaf_parent.alloc(T)); CU(leaf_box.alloc(T)); CU(node_box.alloc(T));
	CU(arrived.alloc(T)); CU(scalars.alloc(4));
	CU(cudaMemset(arrived.p, 0, sizeof(unsigned int) * T)); CU(cudaMemset(scalars.p, 0, 4 * sizeof(unsigned int)));
	BvhNode* out_nodes = nullptr; BvhTri* out_tris = nullptr; Qbvh4Node* wide = nullptr;
	CU(cudaMalloc(&out_nodes, sizeof(BvhNode) * (size_t) (T - 1)));
	CU(cudaMalloc(&out_tris, sizeof(BvhTri) * (size_t) T));
	CU(cudaMemset(out_nodes, 0, sizeof(BvhNode) * (size_t) (T - 1)));   // slots of nodes inside leaves stay unreferenced
	const int block = 256, grid = (T + block - 1) / block;
	CU(cudaEventRecord(ev[0]));
	keys_kernel<<<grid, block>>>(positions, (uint32_t) T, keys.p, values.p);
	size_t temp_bytes = 0;
	CU(cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, keys.p, keys_sorted.p, values.p, order.p, T, 0, 63));
	CU(temp.alloc(temp_bytes));
	CU(cub::DeviceRadixSort::SortPairs(temp.p, temp_bytes, keys.p, keys_sorted.p, values.p, order.p, T, 0, 63));
	hierarchy_kernel<<<grid, block>>>(keys_sorted.p, T, children.p, range.p, inner_parent.p, leaf_parent.p);
	CU(cudaEventRecord(ev[1]));
	tris_kernel<<<grid, block>>>(positions, order.p, (uint32_t) T, dq, out_tris, leaf_box.p);
	boxes_kernel<<<grid, block>>>(T, children.p, inner_parent.p, leaf_parent.p, leaf_box.p, node_box.p, arrived.p);
	depth_kernel<<<grid, block>>>(T, inner_parent.p, leaf_parent.p, scalars.p);
	nodes_kernel<<<grid, block>>>(T, max_leaf, pad, children.p, range.p, leaf_box.p, node_box.p, out_nodes);
	CU(cudaEventRecord(ev[2]));
	// 4-wide collapse: frontier of (binary node, output slot), one launch per level
	CU(cudaMalloc(&wide, sizeof(Qbvh4Node) * (size_t) T));
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
	Qbvh4Node* compact = nullptr;
	CU(cudaMalloc(&compact, sizeof(Qbvh4Node) * (size_t) host_scalars[2]));
	CU(cudaMemcpy(compact, wide, sizeof(Qbvh4Node) * (size_t) host_scalars[2], cudaMemcpyDeviceToDevice));
	cudaFree(wide);
	for (int i = 0; i != 3; ++i) CU(cudaEventElapsedTime(&ms[i], ev[i], ev[i + 1]));
	for (auto& e : ev) cudaEventDestroy(e);
	*nodes = out_nodes; *tris = out_tris; *nodes4 = compact;
	counts[0] = (uint64_t) (T - 1); counts[1] = host_scalars[2];
	depths[0] = host_scalars[0]; depths[1] = levels;
	return 0;
}
