/* ref_host_stubs.c -- lets the reference's camera.c link without GLFW (its control_camera() is never
 * called) and exports the header-only helpers of math_utilities.h. TEST INFRASTRUCTURE. */
#include <stdint.h>
#include "math_utilities.h"   /* found through -I /root/reference/src */

typedef struct GLFWwindow GLFWwindow;
int glfwGetKey(GLFWwindow* w, int k) { (void) w; (void) k; return 0; }
int glfwGetMouseButton(GLFWwindow* w, int b) { (void) w; (void) b; return 0; }
void glfwGetCursorPos(GLFWwindow* w, double* x, double* y) { (void) w; *x = 0.0; *y = 0.0; }
double glfwGetTime(void) { return 0.0; }

void ref_matrix_inverse(float inverse[4][4], const float matrix[4][4]) { matrix_inverse(inverse, matrix); }
uint32_t ref_wang_random_number(uint32_t seed) { return wang_random_number(seed); }
float ref_half_to_float(uint16_t half) { return half_to_float(half); }
