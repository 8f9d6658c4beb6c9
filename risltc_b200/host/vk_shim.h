/* vk_shim.h -- the handful of Vulkan names that the kept host structs mention
 * (scene.h:47-166, ltc_table.h:41-57, main.h:123-156, vulkan_basics.h), defined without
 * Vulkan. Handle slots carry CUDA-side objects instead: buffers and images hold the
 * host staging copy that was uploaded through the C ABI, device_t holds the
 * risltc_device_t. Nothing here talks to a driver. */
#ifndef RISLTC_VK_SHIM_H
#define RISLTC_VK_SHIM_H
#include <stdint.h>
#include <stddef.h>

typedef uint32_t VkBool32;
typedef uint64_t VkDeviceSize;
#define VK_TRUE 1u
#define VK_FALSE 0u

typedef struct VkExtent2D { uint32_t width, height; } VkExtent2D;

/* Only the formats that occur on this path: .vkt payloads (tools/texture_conversion/main.c:32-38),
 * the LTC arrays (ltc_table.c:126,142) and the mesh views (scene.c:80-83). */
typedef enum VkFormat {
	VK_FORMAT_UNDEFINED = 0,
	VK_FORMAT_R8_UINT = 13,
	VK_FORMAT_R8G8_SINT = 21,
	VK_FORMAT_R16G16_UNORM = 77,
	VK_FORMAT_R16G16B16_SFLOAT = 90,
	VK_FORMAT_R16G16B16A16_UNORM = 91,
	VK_FORMAT_R16G16B16A16_SFLOAT = 97,
	VK_FORMAT_R32G32_UINT = 101,
	VK_FORMAT_R32G32B32_SFLOAT = 106,
	VK_FORMAT_R32G32B32A32_SFLOAT = 109,
	VK_FORMAT_BC1_RGB_UNORM_BLOCK = 131,
	VK_FORMAT_BC1_RGB_SRGB_BLOCK = 132,
	VK_FORMAT_BC5_UNORM_BLOCK = 141
} VkFormat;

/* Opaque 64-bit handles. A non-zero value means "resident on the CUDA device". */
typedef uint64_t VkBuffer;
typedef uint64_t VkBufferView;
typedef uint64_t VkDeviceMemory;
typedef uint64_t VkSampler;
typedef uint64_t VkImage;
typedef uint64_t VkImageView;
typedef uint64_t VkAccelerationStructureKHR;

struct risltc_device_s;

/* device_t (vulkan_basics.h:28-77), reduced to what this path needs. */
typedef struct device_s {
	/* The CUDA device object of include/risltc_cuda.h */
	struct risltc_device_s* cuda;
	/* CUDA ordinal it was created for */
	int cuda_ordinal;
	/* Always VK_TRUE: shadow rays are traced by the software BVH kernels */
	VkBool32 ray_tracing_supported;
} device_t;

/* buffer_t / buffers_t (vulkan_basics.h): one sub-allocation and a group of them. `memory` of the
 * group points at the host staging copy (malloc), which is what load_scene keeps after upload. */
typedef struct buffer_s {
	VkBuffer buffer;
	VkDeviceSize offset, size;
} buffer_t;

typedef struct buffers_s {
	buffer_t* buffers;
	uint32_t buffer_count;
	VkDeviceMemory memory;
	VkDeviceSize size;
} buffers_t;

typedef struct image_s {
	VkImage image;
	VkImageView view;
	VkFormat format;
	uint32_t width, height, layers;
	/* Host copy (malloc), uploaded through the C ABI: all layers of an array texture, or every mip level of a material
	 * texture, largest first, decoded to RGBA texels of `texel_format` (RISLTC_TEXEL_*, include/risltc_cuda.h) */
	void* host_data;
	size_t host_size;
	uint32_t mip_count, texel_format;
} image_t;

typedef struct images_s {
	image_t* images;
	uint32_t image_count;
	VkDeviceMemory memory;
} images_t;

#endif
