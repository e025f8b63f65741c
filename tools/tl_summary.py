import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "events ms/step", round(d["decode_ms_per_step_events"], 3), "timeline us/step", round(d["timeline_us_per_step"], 1),
      "gemv us/step", round(d["gemv_us_per_step"], 1), "gemv GB/s", round(d["gemv_gbs"] * 1e0, 1))
for k, v in d["rows"].items():
    print("    %-34s %8.2f us  %5.1f%%" % (k, v["us_each"], v["share"] * 100))
print("    attention CTA phases (ns): entry->prefetch issued, ->pdl_wait done, ->Q staged, ->KV loop done, ->partials written:", d.get("attn_phases_ns"))
