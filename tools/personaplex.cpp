// personaplex — the reference's tools/personaplex.cpp entry point: the speech-to-speech loop with a voice prompt (-v), a text
// system prompt (-p) and the PersonaPlex prompt replay inside moshi_lm_start (src/moshi/models/lm.h:983-1134).
#define MOSHI_TOOL_NO_MAIN
#include "moshi-sts.cpp"

int main(int argc, char **argv) { return sts_main(argc, argv, true); }
