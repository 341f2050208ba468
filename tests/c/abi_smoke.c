#include <stdio.h>
#include "diinn_b200.h"
#include "diinn_b200_debug.h"
int main(void) {
  diinn_config cfg = {64, 256, 4, 3, 0, 0};
  diinn_handle* h = NULL;
  int rc = diinn_create(&h, &cfg);
  printf("%s rc=%d err=%s\n", diinn_version(), rc, diinn_last_error(NULL));
  diinn_output_transform t = {1, 0.5f, 0.5f, 1, 0.f, 1.f, 0};
  (void)t;
  if (h) diinn_destroy(h);
  return rc == DIINN_OK || rc == DIINN_ERR_UNSUPPORTED_DEVICE ? 0 : 1;
}
