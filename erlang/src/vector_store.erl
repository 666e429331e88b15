%% vector_store -- GPU-backed replacement for ErlVectorDB's store process.
%%
%% Same registered-process-per-store model, call protocol and replies as the
%% reference module (its src/vector_store.erl:16-19,38-57): insert/3, search/3,
%% delete/2, get_stats/1, sync/1, get_all_vectors/1, all gen_server:call/2.
%% What changes is where the vectors live: the process keeps only
%%     ids    :: #{Id => Slot}          slots :: #{Slot => Id}
%%     meta   :: #{Id => Metadata}      dev   :: evdb_nif resource (device columns)
%% and the distance scan + top-k + exact re-rank run in libevdb_b200.
%% The final {Distance, Id} ordering among exact ties is re-applied here with
%% lists:sort/1, exactly the reference's tie-break (its perform_search/3).
%%
%% Additive API: search/4 and search_batch/4 take #{metric => cosine | euclidean
%% | manhattan}; the 3-ary forms keep the reference behaviour (cosine).
%% NOTE: written for this repo from the behaviour documented in SURVEY.md; it
%% cannot be compiled here (no Erlang/OTP in the image).
-module(vector_store).
-behaviour(gen_server).

-export([start_link/1, insert/3, search/3, search/4, search_batch/4, delete/2,
         get_stats/1, sync/1, get_all_vectors/1]).
-export([init/1, handle_call/3, handle_cast/2, handle_info/2, terminate/2, code_change/3]).

-record(st, {name, dev, dim = undefined, ids = #{}, slots = #{}, meta = #{},
             ordered = true, persistence = false}).

start_link(Name) -> gen_server:start_link({local, Name}, ?MODULE, [Name], []).
insert(Store, Id, Data) -> gen_server:call(Store, {insert, Id, Data}).
search(Store, Q, K) -> gen_server:call(Store, {search, Q, K, cosine}).
search(Store, Q, K, Opts) -> gen_server:call(Store, {search, Q, K, maps:get(metric, Opts, cosine)}).
search_batch(Store, Qs, K, Opts) ->
    gen_server:call(Store, {search_batch, Qs, K, maps:get(metric, Opts, cosine)}).
delete(Store, Id) -> gen_server:call(Store, {delete, Id}).
get_stats(Store) -> gen_server:call(Store, get_stats).
sync(Store) -> gen_server:call(Store, sync).
get_all_vectors(Store) -> gen_server:call(Store, get_all_vectors).

init([Name]) ->
    process_flag(trap_exit, true),
    Device = application:get_env(erlvectordb, gpu_device, 0),
    Dtype = dtype_code(application:get_env(erlvectordb, gpu_dtype, f32)),
    {ok, Dev} = evdb_nif:new(Device, Dtype, 1),
    Persist = application:get_env(erlvectordb, persistence_enabled, true),
    S0 = #st{name = Name, dev = Dev, persistence = Persist},
    case Persist of
        false -> {ok, S0};
        true ->
            {ok, _} = vector_persistence:start_link(Name),
            case vector_persistence:load_vectors(Name) of
                {ok, Loaded} when map_size(Loaded) > 0 -> {ok, bulk_load(Loaded, S0)};
                _ -> {ok, S0}
            end
    end.

%% one host->device copy of N x D instead of N list traversals
bulk_load(Loaded, S) ->
    Ids = maps:keys(Loaded),
    D = length(maps:get(vector, maps:get(hd(Ids), Loaded))),
    Bin = << <<(float(X)):32/float-native>> || Id <- Ids, X <- maps:get(vector, maps:get(Id, Loaded)) >>,
    ok = evdb_nif:bulk_load(S#st.dev, Bin, length(Ids), D),
    Numbered = lists:zip(Ids, lists:seq(0, length(Ids) - 1)),
    S#st{dim = D,
         ids = maps:from_list(Numbered),
         slots = maps:from_list([{Sl, Id} || {Id, Sl} <- Numbered]),
         meta = maps:map(fun(_, V) -> maps:get(metadata, V) end, Loaded),
         ordered = Ids =:= lists:sort(Ids)}.

handle_call({insert, Id, #{vector := V, metadata := M}}, _From, S) ->
    case check(V, S#st.dim) of
        {error, _} = E -> {reply, E, S};
        {ok, D} ->
            {Slot, New} = case maps:find(Id, S#st.ids) of
                              {ok, Sl} -> {Sl, false};
                              error -> {map_size(S#st.ids), true}
                          end,
            case evdb_nif:upsert(S#st.dev, Slot, V) of
                ok ->
                    Ordered = S#st.ordered andalso
                        (not New orelse Slot =:= 0 orelse maps:get(Slot - 1, S#st.slots) < Id),
                    S1 = S#st{dim = D, ordered = Ordered,
                              ids = maps:put(Id, Slot, S#st.ids),
                              slots = maps:put(Slot, Id, S#st.slots),
                              meta = maps:put(Id, M, S#st.meta)},
                    S#st.persistence andalso vector_persistence:save_vector(S#st.name, Id, V, M),
                    {reply, ok, S1};
                {error, _} = E -> {reply, E, S}
            end
    end;
handle_call({search, Q, K, Metric}, _From, S) ->
    case check(Q, S#st.dim) of
        {error, _} = E -> {reply, E, S};
        {ok, _} when map_size(S#st.ids) =:= 0 -> _ = lists:sublist([], K), {reply, {ok, []}, S};
        {ok, _} -> {reply, {ok, ranked(Q, K, Metric, S)}, S}
    end;
handle_call({search_batch, Qs, K, Metric}, _From, S) ->
    case [E || Q <- Qs, {error, _} = E <- [check(Q, S#st.dim)]] of
        [E | _] -> {reply, E, S};
        [] -> {reply, {ok, [ranked(Q, K, Metric, S) || Q <- Qs]}, S}  % one NIF call per batch in production
    end;
handle_call({delete, Id}, _From, S) ->
    case maps:take(Id, S#st.ids) of
        error -> {reply, ok, S};
        {Slot, Ids1} ->
            {ok, Moved} = evdb_nif:delete(S#st.dev, Slot),
            S1 = case Moved of
                     none -> S#st{ids = Ids1, slots = maps:remove(Slot, S#st.slots)};
                     From ->
                         MovedId = maps:get(From, S#st.slots),
                         S#st{ids = maps:put(MovedId, Slot, Ids1), ordered = false,
                              slots = maps:put(Slot, MovedId, maps:remove(From, S#st.slots))}
                 end,
            S#st.persistence andalso vector_persistence:delete_vector(S#st.name, Id),
            {reply, ok, S1#st{meta = maps:remove(Id, S#st.meta)}}
    end;
handle_call(get_stats, _From, S) ->
    {reply, {ok, #{name => S#st.name, count => map_size(S#st.ids), dimension => S#st.dim,
                   persistence_enabled => S#st.persistence}}, S};
handle_call(sync, _From, #st{persistence = false} = S) -> {reply, {error, persistence_disabled}, S};
handle_call(sync, _From, S) -> {reply, vector_persistence:sync(S#st.name), S};
handle_call(get_all_vectors, _From, S) ->
    All = maps:map(fun(Id, Slot) ->
                           {ok, V} = evdb_nif:get(S#st.dev, Slot, S#st.dim),
                           #{vector => V, metadata => maps:get(Id, S#st.meta)}
                   end, S#st.ids),
    {reply, {ok, All}, S};
handle_call(_Other, _From, S) -> {reply, {error, unknown_request}, S}.

handle_cast(_, S) -> {noreply, S}.
handle_info(_, S) -> {noreply, S}.
terminate(_Reason, S) ->
    S#st.persistence andalso vector_persistence:close_store(S#st.name),
    ok.  % the NIF resource destructor frees device memory when this process dies
code_change(_Old, S, _Extra) -> {ok, S}.

%% ---- internals -----------------------------------------------------------------
check(V, _) when not is_list(V) -> {error, invalid_vector_format};
check(V, Dim) ->
    case lists:all(fun is_number/1, V) of
        false -> {error, invalid_vector_format};
        true when Dim =:= undefined -> {ok, length(V)};
        true when length(V) =:= Dim -> {ok, Dim};
        true -> {error, dimension_mismatch}
    end.

%% Device order is (Distance, Slot).  When slot order is not Id order, ask for a wider window so
%% that every member of the tie group straddling position K is present, then re-sort on
%% {Distance, Id} -- the reference's lists:sort/1 order.
ranked(Q, K, Metric, #st{ordered = true} = S) ->
    {ok, Hits} = evdb_nif:search(S#st.dev, Q, K, Metric),
    [{Id, maps:get(Id, S#st.meta), D} || {D, Slot} <- Hits, Id <- [maps:get(Slot, S#st.slots)]];
ranked(Q, K, Metric, S) -> ranked_wide(Q, K, Metric, S, K + 16).

ranked_wide(Q, K, Metric, S, K2) ->
    N = map_size(S#st.ids),
    {ok, Hits} = evdb_nif:search(S#st.dev, Q, min(K2, N), Metric),
    Tied = K > 0 andalso length(Hits) > K andalso K2 < N andalso
        element(1, lists:nth(K, Hits)) == element(1, lists:last(Hits)),
    case Tied of
        true -> ranked_wide(Q, K, Metric, S, K2 * 2);
        false ->
            Sorted = lists:sort([{D, maps:get(Slot, S#st.slots)} || {D, Slot} <- Hits]),
            [{Id, maps:get(Id, S#st.meta), D} || {D, Id} <- lists:sublist(Sorted, K)]
    end.

dtype_code(f32) -> 0;
dtype_code(bf16) -> 1;
dtype_code(quantization_8bit) -> 2;
dtype_code(quantization_4bit) -> 3.
