// mock: lib/actions/ferm/invert/multi_syssolver_mdagm_aggregate.h
