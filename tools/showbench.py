import json,sys
l=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('value',round(l["value"]), 'e2e',round(l["e2e"]["value"]), 'ms/step',round(l['ms_per_step'],2), 'cpu',l.get("cpu_baseline",{}).get('value'), l.get("parity_check"), 'clocks',l['clocks'])
for e in l["extra"]["configs"]:
  if e["config"] == "single_call": print("single_call gpu wall", e["gpu_call_wall_ms"], "kernel", e["gpu_kernel_ms"], "cpu 1 thread", e["cpu_single_thread_ms"]); continue
  print(e["config"], e.get("N_hor"), e.get("Nobs"), e["batch_per_gpu"], round(e["value"]), e["unit"], "ms", round(e["ms_per_step"],1), e["exit_status_counts"], "gen_s", round(e["workload_generation_s"],1), "it", round(e["inner_iterations_mean"]))
