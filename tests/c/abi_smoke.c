/* Plain-C consumer of include/mr_b200.h: proves the header is valid C99 and that the host-only
 * entry points work without a GPU (run by tests/test_capi_cpu.py). */
#include <stdio.h>
#include <string.h>

#include "mr_b200.h"

int main(int argc, char **argv) {
  const char *bad = "not a model";
  mr_model_info info;
  mr_status s = mr_model_inspect(MR_BOOSTER_LIGHTGBM, (const uint8_t *)bad, strlen(bad), 0, &info);
  if (s != MR_ERR_PARSE) { printf("expected MR_ERR_PARSE, got %d\n", (int)s); return 1; }
  if (strstr(mr_last_error(), "max_feature_idx") == NULL) { printf("unexpected message: %s\n", mr_last_error()); return 1; }
  if (mr_hash64("p1", 2) == 0 || mr_hash64("p1", 2) != mr_hash64("p1", 2) || mr_hash64("p1", 2) == mr_hash64("p2", 2)) return 2;
  if (mr_token_count("foo, bar, baz!", 14) != 3 || mr_token_count(" lead", 5) != 2 || mr_token_count("", 0) != 1) return 3;
  {
    const char *schema = "{\"features\":[{\"name\":\"price\",\"type\":\"number\",\"scope\":\"item\",\"source\":\"item.price\"},"
                         "{\"name\":\"ctr\",\"type\":\"rate\",\"top\":\"click\",\"bottom\":\"impression\",\"bucket\":\"24h\",\"periods\":[7,30]}],"
                         "\"model_features\":[\"ctr\",\"price\"]}";
    mr_schema *sc = NULL;
    int32_t dim = 0;
    if (mr_schema_create(NULL, schema, strlen(schema), &sc) != MR_OK) { printf("schema: %s\n", mr_last_error()); return 4; }
    if (mr_schema_dim(sc) != 3 || mr_schema_feature_offset(sc, "price", &dim) != 2 || dim != 1) return 5;
    if (mr_schema_feature_offset(sc, "ctr", &dim) != 0 || dim != 2) return 6;
    mr_schema_free(sc);
  }
  if (argc > 1 && strcmp(argv[1], "--expect-no-gpu") == 0) {
    mr_ctx *ctx = NULL;
    s = mr_init(0, &ctx);
    if (s != MR_ERR_NO_DEVICE || ctx != NULL) { printf("expected MR_ERR_NO_DEVICE, got %d\n", (int)s); return 7; }
  }
  printf("%s ok\n", mr_version());
  return 0;
}
