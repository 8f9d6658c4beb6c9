/* screenshot.c -- implement_screenshot (main.c:2358-2409) for the offline path: the copy pass (copy_pass.frag.glsl)
 * turns the accumulated frame into 8-bit frames on the device, the host stores them.
 *   *.png  the displayed image (linear -> sRGB, 8 bit)
 *   *.hdr  two 8-bit frames holding the low and the high byte of every channel's half-float bits are combined into
 *          floats (combine_ldr_screenshots_into_hdr, main.c:2339-2350) and stored as Radiance RGBE
 * The reference writes both through the vendored stb_image_write.h; the encoders here are written from the formats'
 * specifications (PNG: RFC 2083 with stored deflate blocks, RFC 1950/1951; Radiance: adaptive run-length scanlines with
 * runs of three or more bytes, literal groups of up to 128, runs of up to 127, which is also what stb emits). */
#include "risltc_host.h"
#include "risltc_cuda.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

float half_to_float(uint16_t h) {
	uint32_t sign = ((uint32_t) h & 0x8000u) << 16, exponent = (h >> 10) & 0x1Fu, mantissa = h & 0x3FFu, x;
	if (exponent == 0) {
		if (mantissa == 0) x = sign;
		else {
			int e = -1;
			do { ++e; mantissa <<= 1; } while (!(mantissa & 0x400u));
			x = sign | ((uint32_t) (127 - 15 - e) << 23) | ((mantissa & 0x3FFu) << 13);
		}
	}
	else if (exponent == 31) x = sign | 0x7F800000u | (mantissa << 13);
	else x = sign | ((exponent + 127 - 15) << 23) | (mantissa << 13);
	float f; memcpy(&f, &x, 4);
	return f;
}

uint16_t float_to_half(float value) {   /* packHalf2x16: round to nearest even */
	uint32_t x; memcpy(&x, &value, 4);
	uint32_t sign = (x >> 16) & 0x8000u, mantissa = x & 0x7FFFFFu;
	int32_t exponent = (int32_t) ((x >> 23) & 0xFFu) - 127 + 15;
	if (((x >> 23) & 0xFFu) == 0xFFu) return (uint16_t) (sign | 0x7C00u | (mantissa ? 0x200u : 0u));
	if (exponent >= 31) return (uint16_t) (sign | 0x7C00u);
	if (exponent <= 0) {
		if (exponent < -10) return (uint16_t) sign;
		mantissa |= 0x800000u;
		uint32_t shift = (uint32_t) (14 - exponent);
		uint32_t half = mantissa >> shift, rest = mantissa & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
		if (rest > halfway || (rest == halfway && (half & 1u))) ++half;
		return (uint16_t) (sign | half);
	}
	uint32_t half = ((uint32_t) exponent << 10) | (mantissa >> 13), rest = mantissa & 0x1FFFu;
	if (rest > 0x1000u || (rest == 0x1000u && (half & 1u))) ++half;
	return (uint16_t) (sign | half);
}

/* main.c:2339-2350: entry i of the HDR image from byte i of the low-bits frame and byte i of the high-bits frame */
void combine_ldr_screenshots_into_hdr(float* hdr, const uint8_t* low_bits, const uint8_t* high_bits, size_t entry_count) {
	for (size_t i = 0; i != entry_count; ++i)
		hdr[i] = half_to_float((uint16_t) (low_bits[i] | ((uint16_t) high_bits[i] << 8)));
}

/* ---------------------------------------------------------------------- *.hdr */

static void rgbe_from_linear(unsigned char rgbe[4], const float* rgb) {
	float largest = fmaxf(rgb[0], fmaxf(rgb[1], rgb[2]));
	if (largest < 1e-32f) { rgbe[0] = rgbe[1] = rgbe[2] = rgbe[3] = 0; return; }
	int exponent;
	float normalize = frexpf(largest, &exponent) * 256.0f / largest;
	rgbe[0] = (unsigned char) (rgb[0] * normalize);
	rgbe[1] = (unsigned char) (rgb[1] * normalize);
	rgbe[2] = (unsigned char) (rgb[2] * normalize);
	rgbe[3] = (unsigned char) (exponent + 128);
}

/* One component plane of a scanline in the adaptive run-length code */
static void write_rle_plane(FILE* file, const unsigned char* plane, uint32_t width) {
	uint32_t x = 0;
	while (x < width) {
		/* the next run of at least three equal bytes starts at r (or nowhere: r = width) */
		uint32_t r = x;
		while (r + 2 < width && !(plane[r] == plane[r + 1] && plane[r] == plane[r + 2])) ++r;
		if (r + 2 >= width) r = width;
		while (x < r) {   /* literals in groups of up to 128 */
			uint32_t length = r - x; if (length > 128) length = 128;
			fputc((int) length, file);
			fwrite(plane + x, 1, length, file);
			x += length;
		}
		if (r + 2 < width) {   /* the run, in pieces of up to 127 */
			while (r < width && plane[r] == plane[x]) ++r;
			while (x < r) {
				uint32_t length = r - x; if (length > 127) length = 127;
				fputc((int) (length + 128), file);
				fputc(plane[x], file);
				x += length;
			}
		}
	}
}

/* rgb: 3 floats per pixel, rows top to bottom */
int write_hdr(const char* path, const float* rgb, uint32_t width, uint32_t height) {
	FILE* file = fopen(path, "wb");
	if (!file) {
		printf("Failed to store a screenshot to the *.hdr file at %s. Please check path and permissions.\n", path);
		return 1;
	}
	fprintf(file, "#?RADIANCE\n# Written by risltc-b200\nFORMAT=32-bit_rle_rgbe\nEXPOSURE=          1.0000000000000\n\n-Y %u +X %u\n", height, width);
	unsigned char* planes = (unsigned char*) malloc((size_t) width * 4);
	for (uint32_t y = 0; y != height; ++y) {
		const float* row = rgb + (size_t) y * width * 3;
		if (width < 8 || width >= 32768) {   /* the format has no run-length code for such scanlines */
			for (uint32_t x = 0; x != width; ++x) {
				unsigned char rgbe[4];
				rgbe_from_linear(rgbe, row + 3 * x);
				fwrite(rgbe, 1, 4, file);
			}
			continue;
		}
		for (uint32_t x = 0; x != width; ++x) {
			unsigned char rgbe[4];
			rgbe_from_linear(rgbe, row + 3 * x);
			for (uint32_t c = 0; c != 4; ++c) planes[x + (size_t) width * c] = rgbe[c];
		}
		const unsigned char header[4] = { 2, 2, (unsigned char) (width >> 8), (unsigned char) (width & 0xFF) };
		fwrite(header, 1, 4, file);
		for (uint32_t c = 0; c != 4; ++c) write_rle_plane(file, planes + (size_t) width * c, width);
	}
	free(planes);
	int failed = ferror(file);
	fclose(file);
	if (failed) { printf("Failed to store a screenshot to the *.hdr file at %s. Please check path and permissions.\n", path); return 1; }
	printf("Wrote screenshot to %s.\n", path);
	return 0;
}

/* Kept entry point: an RGBA32F frame to *.hdr; the values pass through fp16 like the two-frame capture does */
int write_hdr_screenshot(const char* path, const float* rgba, uint32_t width, uint32_t height) {
	size_t pixels = (size_t) width * height;
	float* rgb = (float*) malloc(pixels * 3 * sizeof(float));
	for (size_t i = 0; i != pixels; ++i)
		for (uint32_t c = 0; c != 3; ++c) rgb[3 * i + c] = half_to_float(float_to_half(rgba[4 * i + c]));
	int result = write_hdr(path, rgb, width, height);
	free(rgb);
	return result;
}

/* ---------------------------------------------------------------------- *.png */

static uint32_t crc32_update(uint32_t crc, const unsigned char* data, size_t size) {
	static uint32_t table[256];
	if (!table[1])
		for (uint32_t n = 0; n != 256; ++n) {
			uint32_t c = n;
			for (int k = 0; k != 8; ++k) c = (c & 1u) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
			table[n] = c;
		}
	for (size_t i = 0; i != size; ++i) crc = table[(crc ^ data[i]) & 0xFFu] ^ (crc >> 8);
	return crc;
}

static void put_be32(unsigned char* p, uint32_t v) { p[0] = (unsigned char) (v >> 24); p[1] = (unsigned char) (v >> 16); p[2] = (unsigned char) (v >> 8); p[3] = (unsigned char) v; }

static void write_chunk(FILE* file, const char type[4], const unsigned char* data, size_t size) {
	unsigned char head[8];
	put_be32(head, (uint32_t) size); memcpy(head + 4, type, 4);
	fwrite(head, 1, 8, file);
	if (size) fwrite(data, 1, size, file);
	uint32_t crc = crc32_update(0xFFFFFFFFu, head + 4, 4);
	crc = crc32_update(crc, data, size) ^ 0xFFFFFFFFu;
	unsigned char tail[4]; put_be32(tail, crc);
	fwrite(tail, 1, 4, file);
}

/* rgb: 3 bytes per pixel, rows top to bottom. 8-bit truecolour, filter 0, zlib stream of stored blocks. */
int write_png(const char* path, const uint8_t* rgb, uint32_t width, uint32_t height) {
	FILE* file = fopen(path, "wb");
	if (!file) {
		printf("Failed to store a screenshot to the *.png file at %s. Please check path and permissions.\n", path);
		return 1;
	}
	const unsigned char signature[8] = { 137, 80, 78, 71, 13, 10, 26, 10 };
	fwrite(signature, 1, 8, file);
	unsigned char ihdr[13];
	put_be32(ihdr, width); put_be32(ihdr + 4, height);
	ihdr[8] = 8; ihdr[9] = 2; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
	write_chunk(file, "IHDR", ihdr, 13);
	/* raw image: every row prefixed by its filter type 0 */
	const size_t row_bytes = (size_t) width * 3 + 1, raw_size = row_bytes * height;
	const size_t block_count = (raw_size + 65534) / 65535;
	const size_t stream_size = 2 + raw_size + 5 * (block_count ? block_count : 1) + 4;
	unsigned char* stream = (unsigned char*) malloc(stream_size);
	unsigned char* raw = (unsigned char*) malloc(raw_size ? raw_size : 1);
	for (uint32_t y = 0; y != height; ++y) {
		raw[row_bytes * y] = 0;
		memcpy(raw + row_bytes * y + 1, rgb + (size_t) y * width * 3, (size_t) width * 3);
	}
	size_t cursor = 0;
	stream[cursor++] = 0x78; stream[cursor++] = 0x01;
	uint32_t adler_a = 1, adler_b = 0;
	size_t done = 0;
	do {
		size_t length = raw_size - done; if (length > 65535) length = 65535;
		stream[cursor++] = (done + length == raw_size) ? 1 : 0;   /* BFINAL, BTYPE = 00 */
		stream[cursor++] = (unsigned char) (length & 0xFF); stream[cursor++] = (unsigned char) (length >> 8);
		stream[cursor++] = (unsigned char) (~length & 0xFF); stream[cursor++] = (unsigned char) ((~length >> 8) & 0xFF);
		memcpy(stream + cursor, raw + done, length);
		for (size_t i = 0; i != length; ++i) { adler_a = (adler_a + raw[done + i]) % 65521u; adler_b = (adler_b + adler_a) % 65521u; }
		cursor += length; done += length;
	} while (done < raw_size);
	put_be32(stream + cursor, (adler_b << 16) | adler_a); cursor += 4;
	write_chunk(file, "IDAT", stream, cursor);
	write_chunk(file, "IEND", NULL, 0);
	free(raw); free(stream);
	int failed = ferror(file);
	fclose(file);
	if (failed) { printf("Failed to store a screenshot to the *.png file at %s. Please check path and permissions.\n", path); return 1; }
	printf("Wrote screenshot to %s.\n", path);
	return 0;
}

/* ------------------------------------------------------------ implement_screenshot */

/* The frame currently in the accumulation buffer to *.png and / or *.hdr (either path may be NULL). The reference needs
 * two presented frames for an HDR screenshot (frame_bits 1 then 2, main.c:2403-2405); offline the same two copy passes
 * run back to back on the same accumulated frame. */
int take_screenshot(application_t* app, const char* path_png, const char* path_hdr) {
	if (!path_png && !path_hdr) return 0;
	if (app->stripe_count != 1) {
		printf("Screenshots are taken from whole frames; this process renders stripe %u of %u.\n", app->stripe_index, app->stripe_count);
		return 1;
	}
	const uint32_t width = app->swapchain.extent.width, height = app->swapchain.extent.height;
	const size_t entry_count = (size_t) width * height * 3;
	uint8_t* ldr_copy = (uint8_t*) malloc(2 * entry_count);
	int result = 0;
	if (path_png) {
		result = risltc_cuda_copy_pass(app->device.cuda, 0, ldr_copy);
		if (!result) result = write_png(path_png, ldr_copy, width, height);
	}
	if (path_hdr && !result) {
		result = risltc_cuda_copy_pass(app->device.cuda, 1, ldr_copy) || risltc_cuda_copy_pass(app->device.cuda, 2, ldr_copy + entry_count);
		if (!result) {
			float* hdr_copy = (float*) malloc(entry_count * sizeof(float));
			combine_ldr_screenshots_into_hdr(hdr_copy, ldr_copy, ldr_copy + entry_count, entry_count);
			result = write_hdr(path_hdr, hdr_copy, width, height);
			free(hdr_copy);
		}
	}
	free(ldr_copy);
	return result;
}
