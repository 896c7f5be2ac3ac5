// TEST FIXTURE, not product code.  The reference's experiments/supercell_kessler_surrogate/inference_ponni.cpp includes
// custom_modules/gather_micro_statistics.h but never uses it (StatisticsGatherer belongs to the surrogate's data-gathering
// workflow, outside the time-stepping hot path).  This empty header stands in for it so that the UNMODIFIED
// driver compiles against miniweatherml_b200/host/*.h in tests/test_host_driver.py.
#pragma once
