/* main.c -- command line of the offline renderer; same flags as the reference's main()
 * (main.c:3044-3078): -run_exp runs every experiment selected by the EXP_* environment variables,
 * -e<N> runs experiment N; -gpu<N> picks the CUDA device. */
#include "risltc_host.h"

int main(int argc, char** argv) {
	return risltc_main(argc, argv);
}
