/* A C99 caller of the C-ABI: compiles against include/rl4rs_b200.h alone, links
 * librl4rs_b200.so, and exercises the entry points that need no GPU -- version, observation widths, the AUGRU kernel
 * rule, and r4_create's argument checking with its error message.  Built and run by tests/test_capi_exports.py. */
#include <stdio.h>
#include <stddef.h>
#include <string.h>
#include "rl4rs_b200.h"

int main(void) {
  r4_config cfg;
  r4_env* env = NULL;
  int rc;
  memset(&cfg, 0, sizeof cfg);
  cfg.env_kind = R4_ENV_SLATE;
  cfg.flags = R4_FLAG_RLLIB_MASK;
  cfg.batch_size = 4; cfg.max_steps = 9; cfg.page_items = 9; cfg.action_size = 284; cfg.action_emb_size = 32;
  cfg.maxlen = 32;                       /* unsupported on purpose: the kernels are built for maxlen = 64 */
  cfg.seq_num = 2; cfg.dense_feature_num = 432; cfg.category_feature_num = 21; cfg.category_hash_size = 1000;
  cfg.emb_size = 128; cfg.hidden_units = 128; cfg.simulator = R4_SIM_DIEN;
  printf("sizeof r4_config %d r4_out %d\n", (int)sizeof(r4_config), (int)sizeof(r4_out));
  printf("abi %d\n", r4_abi_version());
  printf("obs_dim dien %d widedeep %d\n", r4_obs_dim(R4_SIM_DIEN), r4_obs_dim(R4_SIM_WIDEDEEP));
  printf("augru_kernel_for 64/148 %d 592/148 %d\n", r4_augru_kernel_for(64, 148), r4_augru_kernel_for(592, 148));
  rc = r4_create(&cfg, 0, &env);
  printf("create rc %d env %s\n", rc, env ? "set" : "null");
  printf("error: %s\n", r4_last_error(NULL));
  return rc == R4_ERR_ARG && env == NULL ? 0 : 1;
}
