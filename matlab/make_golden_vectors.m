function make_golden_vectors(repo_root, reference_root)
%MAKE_GOLDEN_VECTORS  Pin the B200 engine's reference-algorithm mode to comm.LDPCDecoder itself.
%
%   Needs MATLAB + Communications Toolbox + a checkout of robmaunder/ldpc-3gpp-matlab (the build image of
%   this repo has none of them, which is why decoder parity is reported as UNPINNED -- DESIGN.md section 3).
%   Run once on a licensed machine:
%
%       cd <repo>/matlab
%       make_golden_vectors('..', '/path/to/ldpc-3gpp-matlab')
%
%   and commit tests/golden/matlab/outputs.mat.  tests/test_matlab_golden.py then compares, when the file is
%   present (it reports itself as SKIPPED otherwise):
%       oracle B (oracle/nrldpc_oracle.c), its numpy twin (oracle/twin.py) and the CUDA NRLDPC_ALG_BP kernel
%       against comm.LDPCDecoder's decisions, NumIterations, FinalParityChecks and soft outputs;
%       the Python mirror of NRLDPCDecoder (algorithm 'Sum-product') against a_hat of the reference chain.
%
%   Part 1 builds the decoder EXACTLY as NRLDPCDecoder.m:120 does --
%       comm.LDPCDecoder('ParityCheckMatrix',H,'MaximumIterationCount',iterations, ...
%                        'IterationTerminationCondition','Parity check satisfied')
%   with H = get_pcm(get_3gpp_base_graph(BG,get_3gpp_set_index(Z)),Z) (NRLDPC.m:433-440) -- and feeds it the
%   committed cw_tilde columns of tests/golden/matlab/inputs.mat (the LLRs of tests/golden/decode_nms.npz in
%   the layout of NRLDPCDecoder.m:262-264: 2Z zeros, LLRs, +Inf filler), one step() per column as at :265.
%   The two extra output ports only add outputs; they do not change the decisions.
%   Part 2 runs the reference's own TX/RX chain (plot_BLER_vs_SNR.m:98-99,118-133) on three parameter sets
%   with rng(0) and stores a, g_tilde and a_hat.

    if nargin < 1, repo_root = '..'; end
    if nargin >= 2, addpath(reference_root); end
    in = load(fullfile(repo_root, 'tests', 'golden', 'matlab', 'inputs.mat'));
    cases = in.cases;
    out = struct('name', {}, 'BG', {}, 'Z', {}, 'iterations', {}, 'stop_hard', {}, 'stop_iters', {}, ...
                 'stop_parity', {}, 'full_hard', {}, 'full_iters', {}, 'stop_soft', {});
    for i = 1:numel(cases)
        c = cases(i);
        BG = c.BG; Z = c.Z; iterations = c.iterations;
        H = get_pcm(get_3gpp_base_graph(BG, get_3gpp_set_index(Z)), Z);        % NRLDPC.m:433-440
        K = size(H,2) - size(H,1);
        batch = size(c.cw_tilde, 2);
        % (a) the reference's configuration (NRLDPCDecoder.m:120) + iteration count and final parity checks
        hStop = comm.LDPCDecoder('ParityCheckMatrix',H,'MaximumIterationCount',iterations, ...
            'IterationTerminationCondition','Parity check satisfied', ...
            'NumIterationsOutputPort',true,'FinalParityChecksOutputPort',true);
        % (b) same decoder without the stop ('Maximum iteration count' is the toolbox default)
        hFull = comm.LDPCDecoder('ParityCheckMatrix',H,'MaximumIterationCount',iterations, ...
            'NumIterationsOutputPort',true);
        % (c) soft output of the whole codeword: the a-posteriori LLRs Q_i behind the decisions of (a)
        hSoft = comm.LDPCDecoder('ParityCheckMatrix',H,'MaximumIterationCount',iterations, ...
            'IterationTerminationCondition','Parity check satisfied', ...
            'DecisionMethod','Soft decision','OutputValue','Whole codeword');
        o.name = c.name; o.BG = BG; o.Z = Z; o.iterations = iterations;
        o.stop_hard = zeros(K, batch, 'uint8'); o.stop_iters = zeros(1, batch); o.stop_parity = zeros(1, batch);
        o.full_hard = zeros(K, batch, 'uint8'); o.full_iters = zeros(1, batch);
        o.stop_soft = zeros(size(H,2), batch);
        for b = 1:batch
            cw_tilde = c.cw_tilde(:, b);                                       % NRLDPCDecoder.m:262-264 layout
            [c_hat, n_it, par] = step(hStop, cw_tilde);                        % :265
            o.stop_hard(:, b) = uint8(c_hat); o.stop_iters(b) = n_it; o.stop_parity(b) = double(~any(par));
            [c_hat, n_it] = step(hFull, cw_tilde);
            o.full_hard(:, b) = uint8(c_hat); o.full_iters(b) = n_it;
            o.stop_soft(:, b) = step(hSoft, cw_tilde);
        end
        out(end+1) = o; %#ok<AGROW>
        fprintf('%s: iterations %s\n', c.name, mat2str(o.stop_iters));
    end

    % ---- part 2: the reference chain end to end (NRLDPCEncoder -> QPSK -> AWGN -> NRDemodulator -> NRLDPCDecoder)
    sets = [20 2 1/5 2.5; 400 2 1/5 -1.0; 3842 2 1/3 1.0; 1000 1 1/3 0.5];  % A, BG, R, Es/N0 dB
    chain = struct('A', {}, 'BG', {}, 'R', {}, 'EsN0', {}, 'G', {}, 'iterations', {}, 'a', {}, 'g_tilde', {}, 'a_hat_empty', {}, 'a_hat', {});
    rng(0);                                                                   % plot_BLER_vs_SNR.m:45
    hMod = NRModulator('Modulation','QPSK'); hDemod = NRDemodulator('Modulation','QPSK');
    hChan = comm.AWGNChannel('NoiseMethod','Signal to noise ratio (SNR)');    % :48-50
    for s = 1:size(sets,1)
        A = sets(s,1); BG = sets(s,2); R = sets(s,3); EsN0 = sets(s,4); iterations = 8;
        G = round(A/R/hMod.Q_m)*hMod.Q_m;                                     % :94
        hEnc = NRLDPCEncoder('A',A,'BG',BG,'G',G,'Q_m',hMod.Q_m);             % :98
        hDec = NRLDPCDecoder('A',A,'BG',BG,'G',G,'Q_m',hMod.Q_m,'I_HARQ',1,'iterations',iterations); % :99
        hChan.SNR = EsN0; hDemod.Variance = 1/10^(EsN0/10);                   % :105-106
        n = 8;
        e.A = A; e.BG = BG; e.R = R; e.EsN0 = EsN0; e.G = G; e.iterations = iterations;
        e.a = zeros(A, n, 'uint8'); e.g_tilde = zeros(G, n); e.a_hat_empty = zeros(1, n); e.a_hat = zeros(A, n, 'uint8');
        for f = 1:n
            a = round(rand(A,1));                                             % :118
            reset(hDec);                                                      % :122
            g = step(hEnc, a); tx = step(hMod, g); rx = step(hChan, tx);      % :129-131
            g_tilde = step(hDemod, rx);                                       % :132
            a_hat = step(hDec, g_tilde);                                      % :133
            e.a(:, f) = uint8(a); e.g_tilde(:, f) = g_tilde; e.a_hat_empty(f) = isempty(a_hat);
            if ~isempty(a_hat), e.a_hat(:, f) = uint8(a_hat); end
        end
        chain(end+1) = e; %#ok<AGROW>
        release(hEnc); release(hDec);
    end

    info.matlab = version; v = ver('comm'); info.comm_toolbox = v.Version; info.date = datestr(now, 30);
    save(fullfile(repo_root, 'tests', 'golden', 'matlab', 'outputs.mat'), 'out', 'chain', 'info', '-v7');
    fprintf('wrote tests/golden/matlab/outputs.mat (MATLAB %s, Communications Toolbox %s)\n', info.matlab, info.comm_toolbox);
end
