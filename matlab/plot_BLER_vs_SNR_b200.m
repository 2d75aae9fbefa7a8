function plot_BLER_vs_SNR_b200(A, R, BG, Modulation, rv_id_sequence, iterations, target_block_errors, target_BLER, EsN0_start, EsN0_delta, seed, batch)
%PLOT_BLER_VS_SNR_B200  The protocol of plot_BLER_vs_SNR.m with the frame loop turned into MATRIX step calls.
%
%   The reference script hands its System objects ONE frame per step (plot_BLER_vs_SNR.m:116-133); a B200 decodes
%   one frame in 0.16 ms but 4096 frames in 3 ms, so a drop-in user reaches the engine's throughput only by
%   handing it (n x batch) matrices.  This script keeps the reference's protocol -- same arguments and defaults
%   (:30-42), G = round(A/R/Q_m)*Q_m (:94), rng(seed) (:45), Es/N0 stepping until BLER <= target_BLER (:104,:169),
%   the HARQ loop over rv_id_sequence (:124-137), the start detection (:139-144), the results file and its
%   "%f\t%e\n" lines (:79-83,:164-166) -- and replaces the per-frame calls by per-batch ones:
%       * comm.LDPCEncoder / comm.LDPCDecoder (NRLDPCEncoder.m:158, NRLDPCDecoder.m:265) -> nrldpc_mex('encode' /
%         'decode', h, matrix): one column per code block;
%       * bit selection / interleaving (NRLDPCEncoder.m:168-225, NRLDPCDecoder.m:172-242) -> index vectors built
%         once per (rv_id) from the NRLDPC getters, applied to whole matrices;
%       * CRC attach / check -> comm.CRCGenerator per column (cheap next to the chain).
%   Only single-code-block transport blocks (C = 1, i.e. A <= 8424 for BG1 / 3824 for BG2) are batched here; for
%   C > 1 use the reference script with B200LDPCDecoder at the NRLDPCDecoder.m:120 seam (INTEGRATION.md).
%
%   UNVERIFIED under MATLAB (none in the build image).  The same protocol, batched the same way, is what
%   ldpc_3gpp_matlab_b200/bler.py runs on device and tests/test_gpu_bler.py checks.
%
%   batch: frames per step call (default 4096).

    if nargin < 1,  A = 3842; end                       % plot_BLER_vs_SNR.m:30-42
    if nargin < 2,  R = 1/3; end
    if nargin < 3,  BG = 2; end
    if nargin < 4,  Modulation = 'QPSK'; end
    if nargin < 5,  rv_id_sequence = 0; end
    if nargin < 6,  iterations = 8; end
    if nargin < 7,  target_block_errors = 3; end
    if nargin < 8,  target_BLER = 1e-3; end
    if nargin < 9,  EsN0_start = 0; end
    if nargin < 10, EsN0_delta = 0.5; end
    if nargin < 11, seed = 0; end
    if nargin < 12, batch = 4096; end

    rng(seed);                                                                     % :45
    hMod = NRModulator('Modulation', Modulation);                                  % :48-50
    hDemod = NRDemodulator('Modulation', Modulation);
    hChan = comm.AWGNChannel('NoiseMethod', 'Signal to noise ratio (SNR)');
    Q_m = hMod.Q_m;
    G = round(A/R/Q_m)*Q_m;                                                        % :94

    try
        p = NRLDPCEncoder('A', A, 'BG', BG, 'G', G, 'Q_m', Q_m);                   % parameters only (NRLDPC getters)
        if p.C ~= 1
            error('ldpc_3gpp_matlab:UnsupportedParameters', 'plot_BLER_vs_SNR_b200 batches single-code-block transport blocks only (C = %d).', p.C);
        end
        Z = p.Z_c; K = p.K; N = p.N; Kp = p.K_prime; B = p.B; E = p.E_r(1);
        hTBCRC = comm.CRCGenerator('Polynomial', p.transport_block_CRC_polynomial);
        hEncB200 = nrldpc_mex('create', BG, Z, iterations, 1);
        hDecB200 = nrldpc_mex('create', BG, Z, iterations, 1);                     % 'Parity check satisfied', NRLDPCDecoder.m:120
        cleanup = onCleanup(@() cellfun(@(h) nrldpc_mex('destroy', h), {hEncB200, hDecB200})); %#ok<NASGU>
    catch ME
        if strcmp(ME.identifier, 'ldpc_3gpp_matlab:UnsupportedParameters')          % :172-176
            warning('ldpc_3gpp_matlab:UnsupportedParameters', 'The requested combination of parameters is not supported. %s', getReport(ME, 'basic', 'hyperlinks', 'on'));
            return
        else
            rethrow(ME);
        end
    end

    % rate-matching index vectors, one set per redundancy version (closed form of the while loops at
    % NRLDPCEncoder.m:187-195 / NRLDPCDecoder.m:226-234: walk the circular buffer from k_0, skip the filler range)
    filler = (max(Kp - 2*Z, 0) + 1 : K - 2*Z).';                                    % 1-based, d-domain (NRLDPCDecoder.m:224)
    sel = cell(numel(rv_id_sequence), 1);
    for v = 1:numel(rv_id_sequence)
        p.rv_id = rv_id_sequence(v);
        N_cb = p.N_cb; k_0 = p.k_0;
        ring = mod(k_0 + (0:N_cb-1).', N_cb) + 1;                                   % one lap of the circular buffer
        ring = ring(~ismember(ring, filler));                                       % filler bits are never sent (:190)
        laps = ceil(E / numel(ring));
        idx = repmat(ring, laps, 1);
        sel{v} = idx(1:E);
    end
    n_lap = numel(ring);
    intl = reshape(reshape((1:E).', E/Q_m, Q_m).', [], 1);                          % f(i + j*Q_m) = e(i*E/Q_m + j), NRLDPCEncoder.m:219-223

    filename = sprintf('results/BLER_vs_SNR_%d_%f_%d_%s_%d_%d_%f_%d_b200.txt', A, R, BG, Modulation, iterations, target_block_errors, EsN0_start, seed);
    fid = fopen(filename, 'w');                                                     % :79-83
    if fid == -1, error('Could not open %s for writing', filename); end

    EsN0 = EsN0_start; BLER = 1; found_start = false;
    while BLER > target_BLER                                                        % :104
        hChan.SNR = EsN0; hDemod.Variance = 1/10^(EsN0/10);                         % :105-106
        block_count = 0; block_error_count = 0; keep_going = true;
        while keep_going && block_error_count < target_block_errors                 % :116
            a = round(rand(A, batch));                                              % :118, one column per frame
            b = zeros(B, batch);
            for i = 1:batch, b(:, i) = step(hTBCRC, a(:, i)); end                   % NRLDPCEncoder.m:70-89
            c = [b; zeros(K - Kp, batch)];                                          % filler encoded as 0 (:153)
            cw = nrldpc_mex('encode', hEncB200, c);                                 % (N+2Z) x batch, NRLDPCEncoder.m:158
            d = cw(2*Z+1:end, :);                                                   % :159-163
            decoded = false(1, batch);
            a_hat = zeros(A, batch);
            d_tilde_buffer = zeros(N, batch);                                       % reset(hDec), :122
            for v = 1:numel(rv_id_sequence)                                         % HARQ loop, :124-137
                e = d(sel{v}, :);                                                   % bit selection
                f = e(intl, :);                                                     % bit interleaving
                tx = step(hMod, f(:));                                              % :130
                rx = step(hChan, tx);                                               % :131
                f_tilde = reshape(step(hDemod, rx), E, batch);                      % :132
                e_tilde = zeros(E, batch); e_tilde(intl, :) = f_tilde;              % NRLDPCDecoder.m:191-195
                d_tilde = zeros(N, batch);
                for lap = 1:ceil(E / n_lap)                                         % repeated bits add (:230); indices within a lap are distinct
                    rows = (lap-1)*n_lap + 1 : min(lap*n_lap, E);
                    d_tilde(sel{v}(rows), :) = d_tilde(sel{v}(rows), :) + e_tilde(rows, :);
                end
                d_tilde_buffer = d_tilde_buffer + d_tilde;                          % I_HARQ = 1, :236-239
                cw_tilde = [zeros(2*Z, batch); d_tilde_buffer];                     % :262
                cw_tilde(2*Z + filler, :) = Inf;                                    % :264
                c_hat = double(nrldpc_mex('decode', hDecB200, cw_tilde, 0));        % K x batch logical, :265
                for i = find(~decoded)                                              % TB CRC, NRLDPCDecoder.m:336-339
                    if isequal(step(hTBCRC, c_hat(1:A, i)), c_hat(1:B, i))
                        decoded(i) = true; a_hat(:, i) = c_hat(1:A, i);
                    end
                end
                if all(decoded), break; end
            end
            frame_error = ~decoded | any(a_hat ~= a, 1);                            % ~isequal(a, a_hat), :146
            if ~found_start && all(frame_error)                                     % start detection, :139-144
                keep_going = false; BLER = 1;
            else
                found_start = true;
                block_error_count = block_error_count + sum(frame_error);
                block_count = block_count + batch;
                BLER = block_error_count / block_count;
            end
        end
        if BLER < 1
            fprintf(fid, '%f\t%e\n', EsN0, BLER);                                   % :164-166
            fprintf('%f\t%e\t(%d blocks)\n', EsN0, BLER, block_count);
        end
        EsN0 = EsN0 + EsN0_delta;                                                   % :169
    end
    fclose(fid);
end
