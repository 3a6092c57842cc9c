// launch_count.h - counts kernel launches issued by the library (bench.py reports it as gpu_launches)
#pragma once
extern "C" unsigned long long chol_launch_counter;
#define CHOL_COUNT_LAUNCH() __sync_fetch_and_add(&chol_launch_counter, 1ull)
