%% evdb_nif -- Erlang face of erlang/c_src/evdb_nif.c (libevdb_b200, include/evdb.h).
%% Every vector-touching call runs on a dirty CPU scheduler (see the NIF table).
-module(evdb_nif).
-export([new/3, upsert/3, bulk_load/4, append/4, bulk_load_codes/6, delete/2, search/4,
         search_batch/6, get/3, stats/1]).
-on_load(init/0).

init() ->
    Dir = case code:priv_dir(erlvectordb) of
              {error, _} -> "priv";
              D -> D
          end,
    %% fails (and keeps the application from starting) when no sm_100 GPU is usable:
    %% there is deliberately no CPU fallback behind this module
    erlang:load_nif(filename:join(Dir, "evdb_nif"), 0).

%% Devices :: [non_neg_integer()] -- one ordinal = a single-device store; several = one handle over N GPUs
new(_Devices, _Dtype, _Shadow) -> erlang:nif_error(nif_not_loaded).
upsert(_Ref, _Slot, _Vector) -> erlang:nif_error(nif_not_loaded).
bulk_load(_Ref, _F32Bin, _N, _D) -> erlang:nif_error(nif_not_loaded).
append(_Ref, _F64Bin, _N, _D) -> erlang:nif_error(nif_not_loaded).
bulk_load_codes(_Ref, _Codes, _Mins, _Scales, _N, _D) -> erlang:nif_error(nif_not_loaded).
delete(_Ref, _Slot) -> erlang:nif_error(nif_not_loaded).
search(_Ref, _Query, _K, _Metric) -> erlang:nif_error(nif_not_loaded).
search_batch(_Ref, _QBin, _B, _D, _K, _Metric) -> erlang:nif_error(nif_not_loaded).
get(_Ref, _Slot, _D) -> erlang:nif_error(nif_not_loaded).
stats(_Ref) -> erlang:nif_error(nif_not_loaded).
