/* features_oracle.c — placeholder translation unit until the CleanupFeatures / HarvestFeatures restatement lands. */
int features_oracle_available(void) { return 0; }
