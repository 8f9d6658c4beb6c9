/* tables_scene.c -- loaders for the two static inputs of the shading path, keeping the
 * reference's entry points and file formats but uploading through the C ABI instead of Vulkan:
 *   load_ltc_table  (ltc_table.c:23-194)  fit<i>.dat -> UNORM16 arrays -> risltc_cuda_upload_ltc
 *   load_scene      (scene.c:409-559)     *.vks + *.vkt -> risltc_cuda_upload_scene / _materials
 * Error behaviour as in the reference: print one line, clean up, return 1. */
#include "risltc_host.h"
#include "risltc_cuda.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

static char* join_strings(const char* a, const char* b, const char* c, const char* d) {
	size_t n = strlen(a) + strlen(b) + strlen(c) + strlen(d) + 1;
	char* r = (char*) malloc(n);
	snprintf(r, n, "%s%s%s%s", a, b, c, d);
	return r;
}

static void destroy_images(images_t* images) {
	if (images->images)
		for (uint32_t i = 0; i != images->image_count; ++i) free(images->images[i].host_data);
	free(images->images);
	memset(images, 0, sizeof(*images));
}

/* ---------------------------------------------------------------- LTC tables */

void destroy_ltc_table(ltc_table_t* table, const device_t* device) {
	(void) device;
	destroy_images(&table->texture_arrays);
	memset(table, 0, sizeof(*table));
}

int load_ltc_table(ltc_table_t* table, const device_t* device, const char* directory, uint32_t fresnel_count) {
	memset(table, 0, sizeof(*table));
	table->fresnel_count = fresnel_count;
	const uint32_t channel_counts[2] = { 4, 2 };
	uint16_t* staging[2] = { NULL, NULL };
	size_t slice_sizes[2] = { 0, 0 };
	for (uint32_t layer = 0; layer != fresnel_count; ++layer) {
		char index_string[16];
		snprintf(index_string, sizeof(index_string), "%u", layer);
		char* file_path = join_strings(directory, "/fit", index_string, ".dat");
		FILE* file = fopen(file_path, "rb");
		if (!file) {
			printf("Failed to open the linearly transformed cosine table at %s.\n", file_path);
			free(file_path); free(staging[0]); free(staging[1]);
			destroy_ltc_table(table, device);
			return 1;
		}
		free(file_path);
		uint64_t resolution = 0;
		if (fread(&resolution, sizeof(resolution), 1, file) != 1 || resolution == 0) {
			printf("The linearly transformed cosine table %u in directory %s is truncated.\n", layer, directory);
			fclose(file); free(staging[0]); free(staging[1]);
			destroy_ltc_table(table, device);
			return 1;
		}
		if (table->roughness_count == 0) {
			table->roughness_count = table->inclination_count = (uint32_t) resolution;
			for (uint32_t j = 0; j != 2; ++j) {
				slice_sizes[j] = (size_t) (resolution * resolution * channel_counts[j]);
				staging[j] = (uint16_t*) malloc(sizeof(uint16_t) * slice_sizes[j] * fresnel_count);
			}
		}
		else if (resolution != table->roughness_count) {
			printf("The linearly transformed cosine tables in directory %s have inconsistent resolutions. One has resolution %llux%llu, another %ux%u.\n",
				directory, (unsigned long long) resolution, (unsigned long long) resolution, table->roughness_count, table->roughness_count);
			fclose(file); free(staging[0]); free(staging[1]);
			destroy_ltc_table(table, device);
			return 1;
		}
		for (uint64_t j = 0; j != resolution * resolution; ++j) {
			float fit[5];
			if (fread(fit, sizeof(float), 5, file) != 5) memset(fit, 0, sizeof(fit));
			/* adjugate of [[a,0,b],[0,c,0],[d,0,1]] (the inverse up to a factor), ltc_table.c:86-90 */
			float inverse[3][3] = {
				{ fit[2], 0.0f, -fit[1] * fit[2] },
				{ 0.0f, fit[0] - fit[1] * fit[3], 0.0f },
				{ -fit[2] * fit[3], 0.0f, fit[0] * fit[2] } };
			/* scale the largest entry to magnitude one so that UNORM16 applies */
			float largest = fabsf(inverse[0][0]);
			for (uint32_t k = 0; k != 3; ++k)
				for (uint32_t l = 0; l != 3; ++l)
					if (largest < fabsf(inverse[k][l])) largest = fabsf(inverse[k][l]);
			for (uint32_t k = 0; k != 3; ++k)
				for (uint32_t l = 0; l != 3; ++l)
					inverse[k][l] /= largest;
			const float entries[6] = { inverse[0][0], inverse[0][2], inverse[1][1], inverse[2][0], inverse[2][2], fit[4] };
			uint32_t entry = 0;
			for (uint32_t k = 0; k != 2; ++k) {
				for (uint32_t l = 0; l != channel_counts[k]; ++l, ++entry) {
					float value = entries[entry];
					value *= (entry == 1) ? -1.0f : 1.0f;
					if (value < 0.0f) value = 0.0f;
					if (value > 1.0f) value = 1.0f;
					staging[k][slice_sizes[k] * layer + channel_counts[k] * j + l] = (uint16_t) (value * 65535.0f + 0.5f);
				}
			}
		}
		fclose(file);
	}
	/* "device local texture arrays": keep the host copies in the images and make them resident */
	table->texture_arrays.image_count = 2;
	table->texture_arrays.images = (image_t*) calloc(2, sizeof(image_t));
	for (uint32_t j = 0; j != 2; ++j) {
		image_t* image = &table->texture_arrays.images[j];
		image->format = (j == 0) ? VK_FORMAT_R16G16B16A16_UNORM : VK_FORMAT_R16G16_UNORM;
		image->width = table->roughness_count; image->height = table->inclination_count; image->layers = fresnel_count;
		image->host_data = staging[j];
		image->host_size = sizeof(uint16_t) * slice_sizes[j] * fresnel_count;
	}
	if (device && device->cuda) {
		if (risltc_cuda_upload_ltc(device->cuda, staging[0], staging[1], table->roughness_count, table->inclination_count, fresnel_count)) {
			printf("Failed to copy linearly transformed cosine coefficients from the staging buffer to device local memory.\n");
			destroy_ltc_table(table, device);
			return 1;
		}
		table->texture_arrays.images[0].image = table->texture_arrays.images[1].image = 1;
		table->sampler = 1;
	}
	/* lookup constants, ltc_table.c:184-191 */
	table->constants.fresnel_index_factor = (float) (table->fresnel_count - 1);
	table->constants.fresnel_index_summand = 0.0f;
	table->constants.roughness_factor = (float) (table->roughness_count - 1) / (float) table->roughness_count;
	table->constants.roughness_summand = 0.5f / (float) table->roughness_count;
	table->constants.inclination_factor = (float) (table->inclination_count - 1) / (0.5f * M_PI_F * table->inclination_count);
	table->constants.inclination_summand = 0.5f / (float) table->inclination_count;
	return 0;
}

/* -------------------------------------------------------------------- scenes */

const char* get_material_texture_suffix(material_texture_type_t type) {
	switch (type) {
	case material_texture_type_base_color: return "BaseColor";
	case material_texture_type_specular: return "Specular";
	case material_texture_type_normal: return "Normal";
	default: return NULL;
	}
}

void destroy_scene(scene_t* scene, const device_t* device) {
	(void) device;
	if (scene->materials.material_names)
		for (uint64_t i = 0; i != scene->materials.material_count; ++i) free(scene->materials.material_names[i]);
	free(scene->materials.material_names);
	destroy_images(&scene->materials.textures);
	free((void*) (uintptr_t) scene->mesh.memory);
	free(scene->acceleration_structure.buffers.buffers);
	memset(scene, 0, sizeof(*scene));
}

/* ---- block-compressed texels. The reference hands BC1 / BC5 blocks to the texture units; here they are decoded once at load
 * time. Block layouts are those of the Vulkan / D3D specification; interpolated palette entries are the exact rational
 * values rounded to the nearest 8-bit value (hardware decoders differ from one another in the last bit here). */
static void decode_bc1_block(const uint8_t block[8], uint8_t rgba[16][4]) {
	const uint32_t c0 = block[0] | ((uint32_t) block[1] << 8), c1 = block[2] | ((uint32_t) block[3] << 8);
	uint32_t palette[4][3];
	const uint32_t endpoints[2] = { c0, c1 };
	for (int e = 0; e != 2; ++e) {
		const uint32_t r = (endpoints[e] >> 11) & 31u, g = (endpoints[e] >> 5) & 63u, b = endpoints[e] & 31u;
		palette[e][0] = (r << 3) | (r >> 2); palette[e][1] = (g << 2) | (g >> 4); palette[e][2] = (b << 3) | (b >> 2);
	}
	for (int k = 0; k != 3; ++k) {
		if (c0 > c1) {
			palette[2][k] = (2u * palette[0][k] + palette[1][k] + 1u) / 3u;
			palette[3][k] = (palette[0][k] + 2u * palette[1][k] + 1u) / 3u;
		}
		else {
			palette[2][k] = (palette[0][k] + palette[1][k] + 1u) / 2u;
			palette[3][k] = 0u;   /* BC1_RGB: the transparent entry reads as opaque black */
		}
	}
	const uint32_t indices = block[4] | ((uint32_t) block[5] << 8) | ((uint32_t) block[6] << 16) | ((uint32_t) block[7] << 24);
	for (int i = 0; i != 16; ++i) {
		const uint32_t* c = palette[(indices >> (2 * i)) & 3u];
		rgba[i][0] = (uint8_t) c[0]; rgba[i][1] = (uint8_t) c[1]; rgba[i][2] = (uint8_t) c[2]; rgba[i][3] = 255;
	}
}

static void decode_bc4_block(const uint8_t block[8], uint8_t values[16]) {
	const uint32_t e0 = block[0], e1 = block[1];
	uint32_t palette[8] = { e0, e1, 0, 0, 0, 0, 0, 0 };
	if (e0 > e1)
		for (uint32_t k = 1; k != 7; ++k) palette[k + 1] = ((7u - k) * e0 + k * e1 + 3u) / 7u;
	else {
		for (uint32_t k = 1; k != 5; ++k) palette[k + 1] = ((5u - k) * e0 + k * e1 + 2u) / 5u;
		palette[6] = 0u; palette[7] = 255u;
	}
	uint64_t indices = 0;
	for (int k = 0; k != 6; ++k) indices |= (uint64_t) block[2 + k] << (8 * k);
	for (int i = 0; i != 16; ++i) values[i] = (uint8_t) palette[(indices >> (3 * i)) & 7u];
}

/* One mip level of width x height texels from BC1 (8-byte blocks) or BC5 (two BC4 blocks: red, green) to RGBA8 */
static void decode_block_compressed_level(uint8_t* rgba, const uint8_t* blocks, uint32_t width, uint32_t height, int bc5) {
	const uint32_t blocks_x = (width + 3) / 4, blocks_y = (height + 3) / 4, block_size = bc5 ? 16u : 8u;
	for (uint32_t by = 0; by != blocks_y; ++by)
		for (uint32_t bx = 0; bx != blocks_x; ++bx) {
			const uint8_t* block = blocks + ((size_t) by * blocks_x + bx) * block_size;
			uint8_t texels[16][4];
			if (bc5) {
				uint8_t red[16], green[16];
				decode_bc4_block(block, red); decode_bc4_block(block + 8, green);
				for (int i = 0; i != 16; ++i) { texels[i][0] = red[i]; texels[i][1] = green[i]; texels[i][2] = 0; texels[i][3] = 255; }
			}
			else decode_bc1_block(block, texels);
			for (uint32_t y = 0; y != 4; ++y)
				for (uint32_t x = 0; x != 4; ++x) {
					const uint32_t px = 4 * bx + x, py = 4 * by + y;
					if (px < width && py < height) memcpy(rgba + 4 * ((size_t) py * width + px), texels[4 * y + x], 4);
				}
		}
}

/* Exposed for tests: decode `width` x `height` texels of BC1 (bc5 == 0) or BC5 (bc5 != 0) blocks to RGBA8 */
void decode_block_compressed_texels(uint8_t* rgba, const uint8_t* blocks, uint32_t width, uint32_t height, int bc5) {
	decode_block_compressed_level(rgba, blocks, width, height, bc5);
}

/* One *.vkt texture (textures.c:95-172, header tools/texture_conversion/main.c:41-63) with its whole mip chain. The formats the
 * texture conversion tool writes are accepted: RGBA32F as it is, BC1 (UNORM / SRGB) and BC5 decoded to RGBA8. */
static int load_vkt_texture(image_t* image, const char* file_path) {
	FILE* file = fopen(file_path, "rb");
	if (!file) {
		printf("Failed to open the texture file at path %s.\n", file_path);
		return 1;
	}
	uint32_t head[6]; uint64_t payload_size = 0;
	if (fread(head, sizeof(uint32_t), 6, file) != 6 || fread(&payload_size, sizeof(uint64_t), 1, file) != 1 || head[0] != 0xbc1bc1 || head[1] != 1) {
		printf("The texture at path %s does not seem to have the correct format. It is supposed to be converted to a custom format for the renderer using the texture conversion utility. Aborting.\n", file_path);
		fclose(file);
		return 1;
	}
	const uint32_t mipmap_count = head[2];
	image->width = head[3]; image->height = head[4]; image->format = (VkFormat) head[5]; image->layers = 1;
	if (mipmap_count == 0 || mipmap_count > 32 || image->width == 0 || image->height == 0) {
		printf("The texture at path %s has an invalid header (%u mipmaps of %ux%u).\n", file_path, mipmap_count, image->width, image->height);
		fclose(file);
		return 1;
	}
	uint32_t resolutions[32][2]; uint64_t sizes[32], offsets[32];
	for (uint32_t k = 0; k != mipmap_count; ++k) {
		uint64_t size_offset[2];
		if (fread(resolutions[k], sizeof(uint32_t), 2, file) != 2 || fread(size_offset, sizeof(uint64_t), 2, file) != 2) { fclose(file); return 1; }
		sizes[k] = size_offset[0]; offsets[k] = size_offset[1];
	}
	uint8_t* payload = (uint8_t*) malloc(payload_size ? payload_size : 1);
	uint32_t eof_marker = 0;
	if (fread(payload, 1, payload_size, file) != payload_size || fread(&eof_marker, sizeof(eof_marker), 1, file) != 1 || eof_marker != 0xE0FE0F) {
		printf("The texture file at path %s seems to be invalid. The texture data is not followed by the expected end of file marker.\n", file_path);
		free(payload); fclose(file);
		return 1;
	}
	fclose(file);
	const int bc1 = image->format == VK_FORMAT_BC1_RGB_UNORM_BLOCK || image->format == VK_FORMAT_BC1_RGB_SRGB_BLOCK, bc5 = image->format == VK_FORMAT_BC5_UNORM_BLOCK;
	if (!bc1 && !bc5 && image->format != VK_FORMAT_R32G32B32A32_SFLOAT) {
		printf("The texture at path %s uses VkFormat %d; supported are RGBA32F (109), BC1 (131, 132) and BC5 (141).\n", file_path, (int) image->format);
		free(payload);
		return 1;
	}
	const size_t texel_size = (bc1 || bc5) ? 4 : 16;
	size_t texel_count = 0;
	for (uint32_t k = 0; k != mipmap_count; ++k) texel_count += (size_t) resolutions[k][0] * resolutions[k][1];
	image->mip_count = mipmap_count;
	image->texel_format = (bc1 || bc5) ? ((image->format == VK_FORMAT_BC1_RGB_SRGB_BLOCK) ? RISLTC_TEXEL_RGBA8_SRGB : RISLTC_TEXEL_RGBA8_UNORM) : RISLTC_TEXEL_RGBA32F;
	image->host_size = texel_count * texel_size;
	image->host_data = malloc(image->host_size ? image->host_size : 1);
	size_t cursor = 0;
	int result = 0;
	for (uint32_t k = 0; k != mipmap_count && !result; ++k) {
		const uint32_t w = resolutions[k][0], h = resolutions[k][1];
		/* the sampler assumes the usual chain: every level half the previous one, rounded down, at least 1 */
		const uint32_t expected_w = (image->width >> k) ? (image->width >> k) : 1, expected_h = (image->height >> k) ? (image->height >> k) : 1;
		const size_t stored = (bc1 || bc5) ? (size_t) ((w + 3) / 4) * ((h + 3) / 4) * (bc5 ? 16 : 8) : (size_t) w * h * 16;
		if (w != expected_w || h != expected_h || offsets[k] + stored > payload_size || sizes[k] < stored) {
			printf("The texture at path %s has an unexpected mipmap %u (%ux%u, %llu bytes).\n", file_path, k, w, h, (unsigned long long) sizes[k]);
			result = 1;
			break;
		}
		uint8_t* level = (uint8_t*) image->host_data + cursor;
		if (bc1 || bc5) decode_block_compressed_level(level, payload + offsets[k], w, h, bc5);
		else memcpy(level, payload + offsets[k], (size_t) w * h * 16);
		cursor += (size_t) w * h * texel_size;
	}
	free(payload);
	if (result) { free(image->host_data); image->host_data = NULL; image->host_size = 0; }
	return result;
}

int load_scene(scene_t* scene, const device_t* device, const char* file_path, const char* texture_path, VkBool32 request_acceleration_structure) {
	memset(scene, 0, sizeof(*scene));
	FILE* file = fopen(file_path, "rb");
	if (!file) {
		printf("Failed to open the scene file at %s.\n", file_path);
		destroy_scene(scene, device);
		return 1;
	}
	uint32_t file_marker = 0, version = 0;
	if (fread(&file_marker, sizeof(file_marker), 1, file) != 1 || fread(&version, sizeof(version), 1, file) != 1) file_marker = 0;
	if (file_marker != 0xabcabc || version != 1) {
		printf("The scene file at path %s is invalid or unsupported. The format marker is 0x%x, the version is %d.\n", file_path, file_marker, version);
		fclose(file);
		destroy_scene(scene, device);
		return 1;
	}
	size_t ok = fread(&scene->materials.material_count, sizeof(uint64_t), 1, file);
	ok += fread(&scene->mesh.triangle_count, sizeof(uint64_t), 1, file);
	ok += fread(scene->mesh.dequantization_factor, sizeof(float), 3, file);
	ok += fread(scene->mesh.dequantization_summand, sizeof(float), 3, file);
	printf("Triangle count: %llu\n", (unsigned long long) scene->mesh.triangle_count);
	if (ok != 8 || scene->mesh.triangle_count == 0) {
		printf("The scene file at path %s is completely empty, i.e. it holds 0 triangles.\n", file_path);
		fclose(file);
		destroy_scene(scene, device);
		return 1;
	}
	scene->materials.material_names = (char**) calloc(scene->materials.material_count ? scene->materials.material_count : 1, sizeof(char*));
	for (uint64_t i = 0; i != scene->materials.material_count; ++i) {
		uint64_t name_length = 0;
		if (fread(&name_length, sizeof(name_length), 1, file) != 1 || name_length > (1u << 20)) { name_length = 0; }
		scene->materials.material_names[i] = (char*) calloc(name_length + 1, 1);
		if (fread(scene->materials.material_names[i], sizeof(char), name_length + 1, file) != name_length + 1) { /* handled by the end marker check */ }
	}
	/* the three mesh buffers, exactly as they go onto the device (scene.c:55-59, :466-468) */
	const uint64_t T = scene->mesh.triangle_count;
	scene->mesh.positions.size = sizeof(uint32_t) * 2 * 3 * T;
	scene->mesh.normals_and_tex_coords.size = sizeof(uint16_t) * 4 * 3 * T;
	scene->mesh.material_indices.size = sizeof(uint8_t) * T;
	scene->mesh.triangle.size = sizeof(int8_t) * 3 * 2;
	VkDeviceSize offset = 0;
	for (uint32_t i = 0; i != mesh_buffer_count_full; ++i) {
		scene->mesh.buffers[i].offset = offset;
		offset += (scene->mesh.buffers[i].size + 15) & ~(VkDeviceSize) 15;
	}
	scene->mesh.size = offset;
	char* staging = (char*) malloc(offset);
	scene->mesh.memory = (VkDeviceMemory) (uintptr_t) staging;
	int truncated = 0;
	for (uint32_t i = 0; i != mesh_buffer_count; ++i)
		if (fread(staging + scene->mesh.buffers[i].offset, scene->mesh.buffers[i].size, 1, file) != 1) truncated = 1;
	const int8_t triangle_vertices[3][2] = { { -1, -1 }, { 3, -1 }, { -1, 3 } };
	memcpy(staging + scene->mesh.triangle.offset, triangle_vertices, sizeof(triangle_vertices));
	uint32_t eof_marker = 0;
	if (fread(&eof_marker, sizeof(eof_marker), 1, file) != 1) eof_marker = 0;
	fclose(file);
	if (truncated || eof_marker != 0xE0FE0F) {
		printf("The scene file at path %s seems to be invalid. The geometry data is not followed by the expected end of file marker.\n", file_path);
		destroy_scene(scene, device);
		return 1;
	}
	/* mesh upload + acceleration structure (the BVH is always built: shadow rays need it) */
	(void) request_acceleration_structure;
	if (device && device->cuda) {
		if (risltc_cuda_upload_scene(device->cuda,
			(const uint32_t*) (staging + scene->mesh.positions.offset),
			(const uint16_t*) (staging + scene->mesh.normals_and_tex_coords.offset),
			(const uint8_t*) (staging + scene->mesh.material_indices.offset),
			T, scene->mesh.dequantization_factor, scene->mesh.dequantization_summand))
		{
			printf("Failed to construct an acceleration structure for the scene file at path %s.\n", file_path);
			destroy_scene(scene, device);
			return 1;
		}
		for (uint32_t i = 0; i != mesh_buffer_count_full; ++i) scene->mesh.buffers[i].buffer = scene->mesh.buffer_views[i] = 1;
		scene->acceleration_structure.bottom_level = scene->acceleration_structure.top_level = 1;
	}
	/* material textures <texture_path>/<material name>_<suffix>.vkt (scene.c:520-543) */
	uint32_t texture_count = (uint32_t) (scene->materials.material_count * material_texture_count);
	scene->materials.textures.image_count = texture_count;
	scene->materials.textures.images = (image_t*) calloc(texture_count ? texture_count : 1, sizeof(image_t));
	float* constants = (float*) calloc(scene->materials.material_count ? scene->materials.material_count : 1, sizeof(float) * 8);
	int result = 0, flat = 1;
	for (uint64_t i = 0; i != scene->materials.material_count && !result; ++i) {
		for (uint32_t j = 0; j != material_texture_count && !result; ++j) {
			char* name = join_strings(texture_path, "/", scene->materials.material_names[i], "_");
			char* path = join_strings(name, get_material_texture_suffix((material_texture_type_t) j), ".vkt", "");
			image_t* image = &scene->materials.textures.images[i * material_texture_count + j];
			result = load_vkt_texture(image, path);
			if (!result && image->texel_format == RISLTC_TEXEL_RGBA32F && image->width * image->height == 1) {
				/* a flat material: the texel every fetch of shading_pass.frag.glsl:630-633 returns */
				const float* texel = (const float*) image->host_data;
				float* c = constants + 8 * i;
				if (j == material_texture_type_base_color) { c[0] = texel[0]; c[1] = texel[1]; c[2] = texel[2]; }
				else if (j == material_texture_type_specular) { c[3] = texel[0]; c[4] = texel[1]; c[5] = texel[2]; }
				else { c[6] = texel[0]; c[7] = texel[1]; }
			}
			else flat = 0;
			free(name); free(path);
		}
	}
	if (!result && device && device->cuda && scene->materials.material_count) {
		/* scenes whose materials are all flat keep the constant path of the kernels; any real texture switches to textureGrad */
		if (flat) result = risltc_cuda_upload_materials(device->cuda, constants, scene->materials.material_count);
		else {
			risltc_texture_t* textures = (risltc_texture_t*) calloc(texture_count, sizeof(risltc_texture_t));
			for (uint32_t i = 0; i != texture_count; ++i) {
				const image_t* image = &scene->materials.textures.images[i];
				textures[i].format = image->texel_format; textures[i].width = image->width; textures[i].height = image->height;
				textures[i].mip_count = image->mip_count; textures[i].texels = image->host_data;
			}
			result = risltc_cuda_upload_textures(device->cuda, textures, texture_count);
			free(textures);
		}
	}
	free(constants);
	if (result) {
		printf("Failed to load material textures for the scene file at path %s using texture path %s.\n", file_path, texture_path);
		destroy_scene(scene, device);
		return 1;
	}
	scene->materials.sampler = 1;
	return 0;
}
